"""Read-only Field / Player / Cell views over ONE env record of the GPU batch, with the getter names the reference's
pygame View reads (src/view/view.py:203-217 and the drawing helpers; src/model/cell.py:234-250, player.py:129-180,
field.py:442-482).  SURVEY §8f rank 4: lets a viewer or a plotting script written against the reference's objects look at
an env that lives on the GPU.  `FieldView(batch.dump(i))` — a snapshot, nothing is written back.

Colours are not part of the env state (the reference draws them from numpy.random when an object is created and never
reads them on the step path): getColor() returns a colour derived from the object's stable id."""
import math


def _color(seed):
    seed = (int(seed) * 2654435761) & 0xFFFFFFFF
    return (60 + seed % 180, 60 + (seed >> 8) % 180, 60 + (seed >> 16) % 180)


class CellView(object):
    def __init__(self, x, y, mass, radius, name="", color=(200, 200, 200), player=None, merge_time=0.0, velocity=(0.0, 0.0)):
        self._x, self._y, self._mass, self._radius = float(x), float(y), float(mass), float(radius)
        self._name, self._color, self._player = name, color, player
        self._merge_time, self._velocity = float(merge_time), velocity

    def getX(self):
        return self._x

    def getY(self):
        return self._y

    def getPos(self):
        return [self._x, self._y]

    def getMass(self):
        return self._mass

    def getRadius(self):
        return self._radius

    def getName(self):
        return self._name

    def getColor(self):
        return self._color

    def getPlayer(self):
        return self._player

    def getMergeTime(self):
        return self._merge_time

    def getVelocity(self):
        return self._velocity

    def isInFov(self, fovPos, fovSize):  # cell.py:169-177
        h = fovSize / 2
        return not (self._x + self._radius < fovPos[0] - h or self._x - self._radius > fovPos[0] + h or
                    self._y + self._radius < fovPos[1] - h or self._y - self._radius > fovPos[1] + h)


class PlayerView(object):
    def __init__(self, rec, k, name):
        p = rec.players[k]
        self._alive = bool(p["alive"])
        self._name = name
        self._fov = (float(p["fov_x"]), float(p["fov_y"]), float(p["fov_size"]))
        self._cmd = (float(p["cmd_x"]), float(p["cmd_y"]))
        self._cells = [CellView(c["x"], c["y"], c["mass"], c["radius"], name, _color(k + 1), self, c["merge_time"],
                                (float(c["svx"]), float(c["svy"])))
                       for c in rec.cells[k][:int(p["n_cells"])]]

    def getCells(self):
        return self._cells

    def getName(self):
        return self._name

    def getIsAlive(self):
        return self._alive

    def getTotalMass(self):  # player.py:129: numpy.sum over the cells' masses
        return float(sum(c.getMass() for c in self._cells)) if self._cells else 0.0

    def getFovPos(self):  # player.py:156-161: mass-weighted centroid, as last computed on the step path
        return [self._fov[0], self._fov[1]]

    def getFovSize(self):
        return self._fov[2]

    def getCommandPoint(self):
        return [self._cmd[0], self._cmd[1]]

    def getSelected(self):
        return False


class FieldView(object):
    def __init__(self, rec):
        import aigar_b200.layout as lay
        L = rec.layout
        self._size = int(L.field_size)
        names = {lay.BOT_NN: "NN", lay.BOT_GREEDY: "Greedy", lay.BOT_RANDOM: "Random"}
        self._players = [PlayerView(rec, k, "%s %d" % (names.get(int(rec.players[k]["bot"]["type"]), "Bot"), k))
                         for k in range(int(L.n_players))]
        # field.pellets: integer pellets by slot, then ex-blob ("fat") pellets by slot — the canonical candidate order
        self._pellets = [CellView(x, y, m, math.sqrt(m / math.pi), color=_color(1000 + s)) for s, x, y, m in rec.pellet_list()]
        self._pellets += [CellView(f["x"], f["y"], f["mass"], f["radius"], color=_color(5000 + s))
                          for s, f in enumerate(rec.fat) if f["mass"] != 0]
        self._viruses = [CellView(v["x"], v["y"], v["mass"], v["radius"], "Virus", (0, 255, 0), None, 0.0,
                                  (float(v["svx"]), float(v["svy"])))
                         for v in rec.viruses[:int(rec.header["n_viruses"][0])]]
        self._blobs = [CellView(b["x"], b["y"], b["mass"], b["radius"], "Blob", _color(9000 + i), None, 0.0,
                                (float(b["svx"]), float(b["svy"])))
                       for i, b in enumerate(rec.blobs[:int(rec.header["n_blobs"][0])])]

    def getWidth(self):
        return self._size

    def getHeight(self):
        return self._size

    def getPellets(self):
        return self._pellets

    def getViruses(self):
        return self._viruses

    def getBlobs(self):
        return self._blobs

    def getPlayers(self):
        return self._players

    def getPlayerCells(self):  # field.py:473-480
        return [c for p in self._players for c in p.getCells()]

    def getTopTenPlayers(self):  # field.py: players by total mass, descending
        return sorted((p for p in self._players if p.getIsAlive()), key=lambda p: p.getTotalMass(), reverse=True)[:10]

    @staticmethod
    def _in_fov(objs, fovPos, fovSize):
        return [o for o in objs if o.isInFov(fovPos, fovSize)]

    def getPelletsInFov(self, fovPos, fovSize):
        return self._in_fov(self._pellets, fovPos, fovSize)

    def getVirusesInFov(self, fovPos, fovSize):
        return self._in_fov(self._viruses, fovPos, fovSize)

    def getBlobsInFov(self, fovPos, fovSize):
        return self._in_fov(self._blobs, fovPos, fovSize)

    def getPlayerCellsInFov(self, fovPos, fovSize):
        return self._in_fov(self.getPlayerCells(), fovPos, fovSize)
