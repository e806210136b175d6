"""Reference harness — TEST INFRASTRUCTURE, never imported by the product.

Runs the UNMODIFIED game code of NILOIDE/A.I.gar (src/model/{model,field,player,cell,bot,
spatialHashTable}.py, imported from /root/reference/src) as the parity oracle.  Only works where
the reference checkout exists (this container); the GPU box uses the golden vectors this module
generated (tools/gen_golden.py -> tests/golden/) and the C restatement in oracle/agar_oracle.c.

Three patches are installed around the reference code, none of which changes its arithmetic
(SURVEY.md §7 step 1):

  (i)   RNG injection: the `numpy` name seen by model.field / model.bot / model.cell is replaced by a
        proxy whose `.random` reads the counter-based Philox streams of oracle/philox.py
        (stream 0 = field.py draws, 1 = bot.py draws, 2 = cell.py colour draws).
  (ii)  canonical candidate order: spatialHashTable.getObjectsFromBuckets (spatialHashTable.py:38-43)
        returns a `set` whose iteration order depends on object addresses; it is replaced by the same
        de-duplicated collection sorted by a stable key (pellets by pool slot, player cells by
        (player, position in player.cells), viruses / blobs by list position).
  (iii) bookkeeping only: creation serials (uid) for player cells, pool slots for pellets, and an
        event recorder wrapped around eat / merge / collide / spawn calls.

The harness also converts the live Python objects into the env RECORD of include/agar_b200.h so that
reference, oracle and GPU states can be compared field by field.
"""
import importlib
import math
import os
import re
import sys
import types

import numpy as _np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from oracle import philox  # noqa: E402
import aigar_b200.layout as lay  # noqa: E402

REF_SRC = os.environ.get("AGAR_REF_SRC", "/root/reference/src")


def reference_available():
    return os.path.isfile(os.path.join(REF_SRC, "model", "field.py"))


_ref = None  # namespace of reference modules once loaded
_current = None  # RefEnv whose streams the shims read


class _RandomProxy(object):
    def __init__(self, stream_name):
        self._name = stream_name

    def _s(self):
        return getattr(_current, self._name)

    def randint(self, lo, hi=None):
        return self._s().randint(lo, hi)

    def random(self):
        return self._s().random()

    def seed(self, *a, **k):  # field.py:32 reseeds from time%pid; neutralised
        pass


class _NumpyProxy(object):
    """Everything forwards to numpy except `.random`."""

    def __init__(self, stream_name):
        self.random = _RandomProxy(stream_name)

    def __getattr__(self, name):
        return getattr(_np, name)


def _canonical_key(obj):
    env = _current
    pl = obj.getPlayer()
    if pl is not None:
        return (1, env.field.players.index(pl), _index_is(pl.cells, obj))
    kind = getattr(obj, "_kind", None)
    if kind == "pellet":
        return (0, obj._fat, obj._slot)
    if kind == "virus":
        return (2, 0, _index_is(env.field.viruses, obj))
    if kind == "blob":
        return (3, 0, _index_is(env.field.blobs, obj))
    raise RuntimeError("object of unknown kind in a hash table: %r" % (obj,))


def _index_is(lst, obj):
    for i, o in enumerate(lst):
        if o is obj:
            return i
    raise ValueError("object not in list")


def load_reference():
    """Import the reference modules once and install the patches."""
    global _ref
    if _ref is not None:
        return _ref
    if not reference_available():
        raise RuntimeError("reference checkout not found at %s" % REF_SRC)
    # model/model.py imports matplotlib and (via rgbGenerator) pygame; neither is on the step path.
    for name in ("matplotlib", "matplotlib.pyplot", "pygame", "pygame.gfxdraw"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["pygame"].gfxdraw = sys.modules["pygame.gfxdraw"]
    sys.path.insert(0, REF_SRC)
    ns = types.SimpleNamespace()
    ns.model = importlib.import_module("model.model")
    ns.field = importlib.import_module("model.field")
    ns.cell = importlib.import_module("model.cell")
    ns.bot = importlib.import_module("model.bot")
    ns.player = importlib.import_module("model.player")
    ns.sht = importlib.import_module("model.spatialHashTable")
    ns.params = importlib.import_module("model.parameters")
    with open(os.path.join(REF_SRC, "model", "networkParameters.py")) as f:
        ns.network_params_src = f.read()
    # (i) RNG injection
    ns.field.numpy = _NumpyProxy("rng_field")
    ns.bot.numpy = _NumpyProxy("rng_bot")
    ns.cell.numpy = _NumpyProxy("rng_colour")

    # (ii) canonical candidate order
    def getObjectsFromBuckets(self, cellIds):
        seen = {}
        for cellId in cellIds:
            for cell in self.buckets[cellId]:
                seen[id(cell)] = cell
        return sorted(seen.values(), key=_canonical_key)

    ns.sht.spatialHashTable.getObjectsFromBuckets = getObjectsFromBuckets

    # (iii') optional obs-grid canonicalisation (AGAR_OBS_CANONICAL): spatialHashTable.py:19 computes
    # cols = int(ceil(size / bucket)); for the FOV grids size / (size / G) sometimes rounds to G + 2^-49 and the
    # table gets G + 1 columns.  Canonical mode rounds the quotient to 9 decimals first.
    orig_sht_init = ns.sht.spatialHashTable.__init__

    def sht_init(self, hashTableSize, bucketSize, left=0, top=0):
        orig_sht_init(self, hashTableSize, bucketSize, left, top)
        if _current is not None and _current.cfg.obs_mode == lay.OBS_CANONICAL:
            self.rows = int(math.ceil(round(hashTableSize / bucketSize, 9)))
            self.cols = self.rows
            self.buckets = {}
            self.clearBuckets()

    ns.sht.spatialHashTable.__init__ = sht_init

    # ... and bins with the mathematically exact floor((pos -+ radius) / gs) instead of int(x / gs) evaluated AT the
    # bucket edges x = bucketLeft + k * gs (spatialHashTable.py:91-112), whose result hangs on the last bit of x / gs.
    from fractions import Fraction
    orig_ids_fp = ns.sht.spatialHashTable.getIdsForAreaFloatingPoint

    def exact_floor_div(v, gs):
        q = math.floor(v / gs)
        fv, fg = Fraction(v), Fraction(gs)
        while q * fg > fv:
            q -= 1
        while (q + 1) * fg <= fv:
            q += 1
        return q

    def ids_fp(self, pos, radius):
        if _current is None or _current.cfg.obs_mode != lay.OBS_CANONICAL:
            return orig_ids_fp(self, pos, radius)
        px, py = pos[0] - self.left, pos[1] - self.top
        gs = self.bucketSize
        ids = set()
        x0, x1 = exact_floor_div(max(0, px - radius), gs), exact_floor_div(min(self.size - 1, px + radius), gs)
        y0, y1 = exact_floor_div(max(0, py - radius), gs), exact_floor_div(min(self.size - 1, py + radius), gs)
        for bx in range(x0, x1 + 1):
            for by in range(y0, y1 + 1):
                ids.add(bx + by * self.cols)
        return ids

    ns.sht.spatialHashTable.getIdsForAreaFloatingPoint = ids_fp

    # (iii) bookkeeping
    Cell, Field = ns.cell.Cell, ns.field.Field
    orig_cell_init = Cell.__init__

    def cell_init(self, x, y, mass, player):
        orig_cell_init(self, x, y, mass, player)
        if player is not None and _current is not None:
            self._uid = _current.next_uid
            _current.next_uid += 1

    Cell.__init__ = cell_init

    orig_split = Cell.split

    def cell_split(self, commandPoint, w, h):
        new = orig_split(self, commandPoint, w, h)
        if self.player is not None:
            env = _current
            env.log(lay.EV_SPLIT, env.field.players.index(self.player), self._uid, new._uid)
        return new

    Cell.split = cell_split

    orig_set_ejecter = Cell.setEjecterCell

    def set_ejecter(self, cell):
        orig_set_ejecter(self, cell)
        env = _current
        self._kind = "blob"
        env.log(lay.EV_EJECT, env.field.players.index(cell.player), cell._uid, _index_is(env.field.blobs, self))

    Cell.setEjecterCell = set_ejecter

    orig_add_pellet = Field.addPellet

    def add_pellet(self, pellet):
        env = _current
        if getattr(pellet, "_kind", None) == "blob":
            slot = env.alloc(env.fat_slots)
            pellet._fat, pellet._slot = 1, slot
            env.fat_slots[slot] = pellet
            env.log(lay.EV_BLOB_TO_PELLET, slot)
        else:
            slot = env.alloc(env.pellet_slots)
            pellet._fat, pellet._slot = 0, slot
            env.pellet_slots[slot] = pellet
            env.log(lay.EV_SPAWN_PELLET, slot, pellet.getX(), pellet.getY(), pellet.getMass())
        pellet._kind = "pellet"
        orig_add_pellet(self, pellet)

    Field.addPellet = add_pellet

    orig_eat_pellet = Field.eatPellet

    def eat_pellet(self, playerCell, pellet):
        env = _current
        env.log(lay.EV_EAT_PELLET, env.field.players.index(playerCell.player), playerCell._uid,
                pellet._slot | (0x10000 if pellet._fat else 0))
        orig_eat_pellet(self, playerCell, pellet)
        (env.fat_slots if pellet._fat else env.pellet_slots)[pellet._slot] = None

    Field.eatPellet = eat_pellet

    orig_eat_blob = Field.eatBlob

    def eat_blob(self, playerCell, blob):
        env = _current
        env.log(lay.EV_EAT_BLOB, env.field.players.index(playerCell.player), playerCell._uid,
                _index_is(self.blobs, blob), blob.getEjecterCell()._uid)
        orig_eat_blob(self, playerCell, blob)

    Field.eatBlob = eat_blob

    orig_eat_virus = Field.eatVirus

    def eat_virus(self, playerCell, virus):
        env = _current
        env.log(lay.EV_EAT_VIRUS, env.field.players.index(playerCell.player), playerCell._uid,
                _index_is(self.viruses, virus), 16 - len(playerCell.player.cells))
        orig_eat_virus(self, playerCell, virus)

    Field.eatVirus = eat_virus

    orig_virus_eat_blob = Field.virusEatBlob

    def virus_eat_blob(self, virus, blob):
        env = _current
        vi, bi, n0 = _index_is(self.viruses, virus), _index_is(self.blobs, blob), len(self.viruses)
        orig_virus_eat_blob(self, virus, blob)
        if len(self.viruses) > n0:
            self.viruses[-1]._kind = "virus"
        env.log(lay.EV_VIRUS_EAT_BLOB, vi, bi, int(len(self.viruses) > n0))

    Field.virusEatBlob = virus_eat_blob

    orig_eat_player_cell = Field.eatPlayerCell

    def eat_player_cell(self, larger, smaller):
        env = _current
        pl = env.field.players
        env.log(lay.EV_EAT_CELL, pl.index(larger.player), larger._uid, pl.index(smaller.player), smaller._uid)
        orig_eat_player_cell(self, larger, smaller)

    Field.eatPlayerCell = eat_player_cell

    orig_delete_player_cell = Field.deletePlayerCell

    def delete_player_cell(self, playerCell):
        env = _current
        player = playerCell.getPlayer()
        orig_delete_player_cell(self, playerCell)
        if not player.getCells():
            env.log(lay.EV_PLAYER_DIED, env.field.players.index(player))
            env.deaths[env.field.players.index(player)] += 1

    Field.deletePlayerCell = delete_player_cell

    orig_merge = Field.mergeCells

    def merge_cells(self, first, second):
        env = _current
        big, small = (first, second) if first.getMass() > second.getMass() else (second, first)
        env.log(lay.EV_MERGE, env.field.players.index(big.player), big._uid, small._uid)
        orig_merge(self, first, second)

    Field.mergeCells = merge_cells

    orig_adjust = Field.adjustCellPositions

    def adjust_positions(self, cell1, cell2, distance, summedRadii):
        env = _current
        env.log(lay.EV_COLLIDE, env.field.players.index(cell1.player), cell1._uid, cell2._uid)
        orig_adjust(self, cell1, cell2, distance, summedRadii)

    Field.adjustCellPositions = adjust_positions

    orig_spawn_virus = Field.spawnVirus

    def spawn_virus(self):
        env = _current
        orig_spawn_virus(self)
        v = self.viruses[-1]
        v._kind = "virus"
        env.log(lay.EV_SPAWN_VIRUS, len(self.viruses) - 1, int(v.getX()), int(v.getY()))

    Field.spawnVirus = spawn_virus

    orig_init_player = Field.initializePlayer

    def init_player(self, player):
        env = _current
        orig_init_player(self, player)
        c = player.cells[0]
        env.log(lay.EV_SPAWN_PLAYER, env.field.players.index(player), c._uid, int(c.getX()), int(c.getY()))

    Field.initializePlayer = init_player

    orig_reset = Field.reset

    def field_reset(self):
        env = _current
        env.pellet_slots = [None] * env.layout.pellet_cap
        env.fat_slots = [None] * env.layout.fat_cap
        orig_reset(self)

    Field.reset = field_reset

    # capture what each NN bot's turn produced (bot.py:195-232)
    Bot = ns.bot.Bot
    orig_get_state = Bot.getStateRepresentation

    def get_state(self):
        st = orig_get_state(self)
        info = _current.turn_info.setdefault(id(self), {})
        info["state"] = st
        info["observed"] = True
        return st

    Bot.getStateRepresentation = get_state
    _ref = ns
    return ns


def make_params(cfg):
    """networkParameters module for a config, built the reference's way: a text-patched copy of
    src/model/networkParameters.py executed as a module (aigar.py:270-298, 797-800)."""
    ns = load_reference()
    n_nn = sum(1 for k in range(cfg.n_players) if cfg.bot_type[k] == lay.BOT_NN)
    n_gr = sum(1 for k in range(cfg.n_players) if cfg.bot_type[k] == lay.BOT_GREEDY)
    n_rd = sum(1 for k in range(cfg.n_players) if cfg.bot_type[k] == lay.BOT_RANDOM)
    subst = {
        "NUM_NN_BOTS": n_nn, "NUM_GREEDY_BOTS": n_gr, "NUM_RANDOM_BOTS": n_rd,
        "VIRUS_SPAWN": bool(cfg.virus_enabled), "ENABLE_SPLIT": bool(cfg.enable_split),
        "ENABLE_EJECT": bool(cfg.enable_eject), "ENABLE_GREEDY_SPLIT": bool(cfg.enable_greedy_split),
        "FRAME_SKIP_RATE": cfg.frame_skip, "GRID_SQUARES_PER_FOV": cfg.grid_squares,
        "REWARD_TERM": cfg.reward_term, "REWARD_SCALE": cfg.reward_scale, "DEATH_TERM": cfg.death_term,
        "DEATH_FACTOR": cfg.death_factor, "MASS_AS_REWARD": bool(cfg.mass_as_reward),
    }
    src = ns.network_params_src
    for name, val in subst.items():
        src, n = re.subn(r"(?m)^(\s*)%s\s*=.*$" % name, lambda m: "%s%s = %r" % (m.group(1), name, val), src, count=1)
        assert n == 1, name
    mod = types.ModuleType("networkParameters_patched")
    exec(compile(src, "networkParameters_patched.py", "exec"), mod.__dict__)
    # explicit channel flags of the config win over the derived defaults
    flags = {"PELLET_GRID": cfg.pellet_grid, "SELF_GRID": cfg.self_grid, "WALL_GRID": cfg.wall_grid,
             "ENEMY_GRID": cfg.enemy_grid, "VIRUS_GRID": cfg.virus_grid, "SELF_GRID_LF": cfg.self_grid_lf,
             "SELF_GRID_SLF": cfg.self_grid_slf, "ENEMY_GRID_LF": cfg.enemy_grid_lf,
             "ENEMY_GRID_SLF": cfg.enemy_grid_slf, "ALL_PLAYER_GRID": cfg.all_player_grid, "USE_FOVSIZE": cfg.use_fovsize,
             "USE_LAST_FOVSIZE": cfg.use_last_fovsize, "USE_TOTALMASS": cfg.use_totalmass,
             "USE_LAST_ACTION": cfg.use_last_action, "USE_SECOND_LAST_ACTION": cfg.use_second_last_action}
    for k, v in flags.items():
        setattr(mod, k, bool(v))
    # the RUN's copy of the flag (the reference's driver rewrites the run folder's networkParameters.py, aigar.py:270-298); the
    # package-global copy that bot.py:402,439 read stays at its stock value
    mod.NORMALIZE_GRID_BY_MAX_MASS = bool(cfg.normalize_grid_by_max_mass)
    mod.NUM_OF_GRIDS = sum(bool(getattr(mod, k)) for k in ("PELLET_GRID", "SELF_GRID", "WALL_GRID", "VIRUS_GRID",
                                                           "ENEMY_GRID", "SIZE_GRID", "SELF_GRID_LF", "SELF_GRID_SLF",
                                                           "ENEMY_GRID_LF", "ENEMY_GRID_SLF", "ALL_PLAYER_GRID"))
    mod.EXTRA_INPUT = (bool(mod.USE_FOVSIZE) + bool(mod.USE_TOTALMASS) + bool(mod.USE_LAST_ACTION) * 4 +
                       bool(mod.USE_SECOND_LAST_ACTION) * 4 + bool(mod.USE_LAST_FOVSIZE))
    mod.STATE_REPR_LEN = mod.GRID_SQUARES_PER_FOV ** 2 * mod.NUM_OF_GRIDS + mod.EXTRA_INPUT
    if cfg.simple_state:  # GRID_VIEW_ENABLED = False: Bot.getSimpleStateRepresentation (12 values).  The reference leaves STATE_REPR_LEN
        # at the grid formula (networkParameters.py:102 does not look at the flag), which is why none of its learners can consume
        # this representation; the record layouts here carry the true length
        mod.GRID_VIEW_ENABLED = False
        mod.STATE_REPR_LEN = lay.SIMPLE_STATE_LEN
    mod.GATHER_EXP = True
    return mod


class _StubAlg(object):
    """Stands in for QLearn/ActorCritic: returns the action the test supplies (bot.py:127,189,206,224)."""
    discrete = False

    def __init__(self, env, agent):
        self.env, self.agent = env, agent

    def __repr__(self):
        return "Stub"

    def reset(self):
        pass

    def decideMove(self, state, updateNoise=True):
        env = self.env
        env.decided[self.agent] = True
        a = env.pending_actions[self.agent]
        return None, [float(v) for v in a[:env.layout.action_len]]


class RefEnv(object):
    """One reference Model driven frame by frame with injected RNG and externally supplied actions."""

    def __init__(self, cfg, seed=0, env_id=0):
        global _current
        ns = load_reference()
        self.cfg = cfg
        self.layout = lay.layout_for_config(cfg)
        self.seed, self.env_id = seed, env_id
        self.rng_field = philox.Stream(seed, env_id, philox.STREAM_FIELD)
        self.rng_bot = philox.Stream(seed, env_id, philox.STREAM_BOT)
        self.rng_colour = philox.Stream(seed, env_id, philox.STREAM_COLOUR)
        self.next_uid = 0
        self.frame = 0
        self.events = []
        self.event_hash = 0xCBF29CE484222325
        self.n_events_frame = 0
        self.pellet_slots = [None] * self.layout.pellet_cap
        self.fat_slots = [None] * self.layout.fat_cap
        self.deaths = [0] * cfg.n_players
        self.turn_info = {}
        self.decided = [False] * self.layout.n_agents
        self.pending_actions = _np.zeros((max(self.layout.n_agents, 1), 4), dtype=_np.float32)
        self.params = make_params(cfg)
        assert self.params.STATE_REPR_LEN == self.layout.state_len, (self.params.STATE_REPR_LEN, self.layout.state_len)
        _current = self
        self.model = ns.model.Model(False, False, self.params)
        self.field = self.model.field
        names = {lay.BOT_NN: "NN", lay.BOT_GREEDY: "Greedy", lay.BOT_RANDOM: "Random"}
        agent = 0
        for k in range(cfg.n_players):
            t = cfg.bot_type[k]
            alg = None
            if t == lay.BOT_NN:
                alg = _StubAlg(self, agent)
                agent += 1
            self.model.createBot(names[t], alg, self.params)
        # draws made while constructing bots (bot.py:93) are not part of the env streams
        self.rng_field.serial = self.rng_bot.serial = 0
        self.next_uid = 0
        self.model.initialize()
        assert self.field.size == self.layout.field_size

    # ---- bookkeeping used by the patches
    def log(self, typ, a=0, b=0, c=0, d=0):
        ev = (int(typ), int(a), int(b), int(c), int(d))
        self.events.append(ev)
        if ev[0] != lay.EV_COLLIDE:  # logged, not hashed: a touching pair's re-test hangs on the last bit
            self.event_hash = lay.event_hash_step(self.event_hash, ev)

    def alloc(self, slots):
        for i, s in enumerate(slots):
            if s is None:
                return i
        raise RuntimeError("pool overflow in the reference harness")

    # ---- driving
    def reset(self):
        global _current
        _current = self
        self.events = []
        self.event_hash = 0xCBF29CE484222325
        self.model.resetModel()
        self.frame = 0

    def reset_bots(self):
        global _current
        _current = self
        self.model.resetBots()

    def step(self, actions=None):
        """One Model.update() (model.py:100-112).  actions: float32 [A][4] used by agents whose
        decideMove() is due this frame.  Returns per-agent turn info."""
        global _current
        _current = self
        if actions is not None:
            self.pending_actions = _np.asarray(actions, dtype=_np.float32).reshape(-1, 4)
        self.events = []
        self.turn_info = {}
        self.decided = [False] * self.layout.n_agents
        nn_bots = self.model.getNNBots()
        n_exp0 = [len(b.experiences) for b in nn_bots]
        self.model.update()
        self.frame += 1
        out = []
        for a, b in enumerate(nn_bots):
            info = self.turn_info.get(id(b), {})
            valid = len(b.experiences) > n_exp0[a]
            done = valid and b.experiences[-1][3] is None
            st = info.get("state")
            out.append({
                "observed": bool(info.get("observed", False)),
                "obs": None if st is None else _np.asarray(st, dtype=_np.float64).reshape(-1),
                "reward": float(b.lastReward),
                "valid": bool(valid), "done": bool(done), "need_action": bool(self.decided[a]),
            })
            self._last_turn = out
        return out

    # ---- conversion to the env record
    def to_record(self, last_turn=None):
        L = self.layout
        rec = lay.Record(L)
        f = self.field
        h = rec.header
        h["rng_field"], h["rng_bot"] = self.rng_field.serial, self.rng_bot.serial
        h["next_uid"], h["frame"] = self.next_uid, self.frame
        h["n_viruses"], h["n_blobs"] = len(f.viruses), len(f.blobs)
        h["n_fat"] = sum(1 for s in self.fat_slots if s is not None)
        h["n_pellets"] = sum(1 for s in self.pellet_slots if s is not None)
        assert h["n_fat"][0] + h["n_pellets"][0] == len(f.pellets)
        h["n_dead"] = len(f.deadPlayers)
        for i, p in enumerate(f.deadPlayers):
            h["dead_order"][0][i] = f.players.index(p)
        h["n_events"] = len(self.events)
        h["event_hash"] = self.event_hash
        in_player_hash = set()
        for b in f.playerHashTable.buckets.values():
            for o in b:
                in_player_hash.add(id(o))
        in_virus_hash = set()
        for b in f.virusHashTable.buckets.values():
            for o in b:
                in_virus_hash.add(id(o))
        bots = self.model.bots
        agent = 0
        for k, pl in enumerate(f.players):
            P = rec.players[k]
            P["alive"], P["respawn_time"], P["n_cells"] = int(pl.isAlive), pl.respawnTime, len(pl.cells)
            P["do_split"], P["do_eject"] = int(bool(pl.doSplit)), int(bool(pl.doEject))
            P["cmd_x"], P["cmd_y"] = pl.commandPoint[0], pl.commandPoint[1]
            valid = pl.fovSize is not None and len(pl.fovPos) == 2
            P["fov_valid"] = int(valid)
            if valid:
                P["fov_x"], P["fov_y"], P["fov_size"] = pl.fovPos[0], pl.fovPos[1], pl.fovSize
            for i, c in enumerate(pl.cells):
                C = rec.cells[k, i]
                C["x"], C["y"], C["mass"], C["radius"] = c.x, c.y, c.mass, c.radius
                C["svx"], C["svy"] = c.splitVelocity[0], c.splitVelocity[1]
                C["merge_time"], C["counter"], C["uid"] = c.mergeTime, c.splitVelocityCounter, c._uid
                C["flags"] = (lay.CF_EJECT if c.blobToBeEjected else 0) | (lay.CF_INHASH if id(c) in in_player_hash else 0)
            b = bots[k]
            B = P["bot"]
            B["type"] = {"NN": lay.BOT_NN, "Greedy": lay.BOT_GREEDY, "Random": lay.BOT_RANDOM}[b.type]
            B["has_action"] = int(b.currentAction is not None)
            if b.currentAction is not None:
                for i, v in enumerate(b.currentAction):
                    B["cur_action"][i] = float(v)
            B["has_last_action"] = int(b.lastAction is not None)
            if b.lastAction is not None:
                for i, v in enumerate(b.lastAction):
                    B["last_action"][i] = float(v)
            B["skip_frames"] = b.skipFrames
            B["has_last_mass"] = int(b.lastMass is not None)
            B["last_mass"] = b.lastMass if b.lastMass is not None else 0.0
            B["has_old_state"] = int(b.oldState is not None)
            B["time"] = b.time
            B["skipping"] = int(bool(b.currentlySkipping))
            B["cum_reward"], B["last_reward"] = b.cumulativeReward, b.lastReward
            B["fov_size_feat"] = b.fovSize if b.fovSize is not None else 0.0
            B["last_fov_size_feat"] = b.lastFovSize if b.lastFovSize is not None else 0.0
            s, mx = 0.0, 0.0
            for m in b.totalMasses:
                s += float(m)
                mx = max(mx, float(m))
            B["stat_mass_sum"], B["stat_mass_max"], B["stat_frames"] = s, mx, len(b.totalMasses)
            B["stat_deaths"] = self.deaths[k]
            if b.type == "NN":
                if last_turn is not None:
                    t = last_turn[agent]
                    B["need_action"], B["exp_valid"], B["exp_done"] = int(t["need_action"]), int(t["valid"]), int(t["done"])
                if L.n_hist:
                    rec.hist[agent, 0] = b.lastSelfGrid
                    rec.hist[agent, 1] = b.secondLastSelfGrid
                    rec.hist[agent, 2] = b.lastEnemyGrid
                    rec.hist[agent, 3] = b.secondLastEnemyGrid
                agent += 1
        for i, v in enumerate(f.viruses):
            V = rec.viruses[i]
            V["x"], V["y"], V["mass"], V["radius"] = v.x, v.y, v.mass, v.radius
            V["svx"], V["svy"], V["counter"] = v.splitVelocity[0], v.splitVelocity[1], v.splitVelocityCounter
            V["aux"] = lay.CF_INHASH if id(v) in in_virus_hash else 0
        for i, bl in enumerate(f.blobs):
            Bm = rec.blobs[i]
            Bm["x"], Bm["y"], Bm["mass"], Bm["radius"] = bl.x, bl.y, bl.mass, bl.radius
            Bm["svx"], Bm["svy"], Bm["counter"] = bl.splitVelocity[0], bl.splitVelocity[1], bl.splitVelocityCounter
            Bm["aux"] = bl.ejecterCell._uid
        for s, p in enumerate(self.pellet_slots):
            if p is not None:
                rec.pellets[s] = lay.pack_pellet(p.x, p.y, p.mass)
        for s, p in enumerate(self.fat_slots):
            if p is not None:
                F = rec.fat[s]
                F["x"], F["y"], F["mass"], F["radius"] = p.x, p.y, p.mass, p.radius
        n = min(len(self.events), L.event_cap)
        for i in range(n):
            rec.events[i] = self.events[i]
        return rec

    # ---- state injection (record -> live reference objects), used for obs-function parity
    def observe_player(self, agent):
        """getStateRepresentation() of NN bot `agent` on the current world without side effects on
        the frame-skip bookkeeping (history grids and fov features DO advance, as in the reference)."""
        global _current
        _current = self
        bot = self.model.getNNBots()[agent]
        st = bot.getStateRepresentation()
        return None if st is None else _np.asarray(st, dtype=_np.float64).reshape(-1)
