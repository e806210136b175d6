"""Known-answer tests captured from the reference (hand-placed scenes, no RNG).
They pin oracle/agar_oracle.c (libm build) to the reference's float64 results wherever it runs."""
import numpy as np
import pytest

import aigar_b200.layout as lay
from oracle import oracle as orc


def scene(cells, pellets=(), cmd=(0.0, 0.0), split=False, eject=False, multi=False, portable=False):
    cfg = lay.derive_config(split=multi, eject=multi, pellet_spawn=False, event_cap=64, overrides={"pellet_cap": 16})
    env = orc.OracleEnv(cfg, seed=0, env_id=0, portable=portable)
    assert env.layout.field_size == 75
    rec = env.record
    rec.pellets[:] = 0
    rec.header["n_pellets"] = len(pellets)
    for s, (x, y, m) in enumerate(pellets):
        rec.pellets[s] = lay.pack_pellet(x, y, m)
    rec.players["n_cells"][0] = len(cells)
    for i, (x, y, m) in enumerate(cells):
        c = rec.cells[0, i]
        c["x"], c["y"], c["mass"], c["radius"] = x, y, m, np.sqrt(m / np.pi)
        c["svx"] = c["svy"] = c["merge_time"] = 0.0
        c["counter"], c["uid"], c["flags"] = 0, i, 0
    rec.header["next_uid"] = len(cells)
    p = rec.players
    p["cmd_x"][0], p["cmd_y"][0], p["do_split"][0], p["do_eject"][0] = cmd[0], cmd[1], int(split), int(eject)
    return env


def test_kat1_move_and_decay():
    env = scene([(10, 10, 10)], cmd=(20, 10))
    env.field_update()
    c = env.record.cells[0, 0]
    assert c["x"] == 11.340207150895663 and c["y"] == 10.0
    assert c["mass"] == 9.996666666666666 and c["radius"] == 1.78382673734978 and c["counter"] == -1


def test_kat2_target_inside_cell():
    env = scene([(10, 10, 10)], cmd=(10.5, 10.5))
    env.field_update()
    c = env.record.cells[0, 0]
    assert c["x"] == 10.148909223515643 and c["y"] == 10.148909223515643


def test_kat3_eat_chain_is_order_dependent():
    # canonical candidate order = slot order: the d=2 pellet is reachable only after the other two grew the cell
    env = scene([(10, 10, 10)], pellets=[(11, 10, 1), (10, 11, 3), (12, 10, 1)], cmd=(10, 10))
    env.field_update()
    assert env.record.cells[0, 0]["mass"] == 14.996666666666666
    assert int(env.record.header["n_pellets"][0]) == 0 and not env.record.pellets.any()
    eaten = [e[3] for e in env.record.event_list() if e[0] == lay.EV_EAT_PELLET]
    assert eaten == [0, 1, 2]
    # ... and is skipped for good when it is visited before the mass-3 pellet (no second pass in the reference)
    env = scene([(10, 10, 10)], pellets=[(11, 10, 1), (12, 10, 1), (10, 11, 3)], cmd=(10, 10))
    env.field_update()
    assert env.record.cells[0, 0]["mass"] == 13.996666666666666
    assert [e[3] for e in env.record.event_list() if e[0] == lay.EV_EAT_PELLET] == [0, 2]


def test_kat4_split():
    env = scene([(30, 30, 100)], cmd=(60, 30), split=True, multi=True)
    env.field_update()
    r = env.record
    assert int(r.players["n_cells"][0]) == 2
    parent, twin = r.cells[0, 0], r.cells[0, 1]
    assert parent["x"] == 30.59864854438848 and parent["mass"] == 49.983333333333334
    assert parent["radius"] == 3.9887578447957712 and parent["merge_time"] == 0
    assert twin["x"] == 32.282047772056615 and twin["svx"] == 2.282047772056613 and twin["svy"] == 0.0
    assert twin["counter"] == 15 and twin["merge_time"] == pytest.approx(392.469175, abs=1e-6)


def test_kat5_eject_leaves_radius_stale():
    env = scene([(30, 30, 100)], cmd=(30, 60), eject=True, multi=True)
    env.field_update()
    r = env.record
    c = r.cells[0, 0]
    assert c["y"] == 30.59864854438848 and c["mass"] == 81.96666666666667 and c["radius"] == 5.640955441132256
    assert int(r.header["n_blobs"][0]) == 1
    b = r.blobs[0]
    assert (b["x"], b["y"]) == (30.0, 30.59864854438848) and b["mass"] == 14.4
    assert b["svx"] == 1.3973512497752397e-16 and b["svy"] == 2.282047772056613 and b["counter"] == 15


def _grid_nonzeros(obs, g=11):
    grid = obs[:g * g].reshape(g, g)
    return sorted((int(r), int(c), float(grid[r, c])) for r, c in zip(*np.nonzero(grid)))


def test_kat6_obs_with_the_ceil_defect():
    env = scene([(37.5, 37.5, 10)], pellets=[(30, 30, 1), (40, 37, 2), (60, 60, 3), (37, 37, 1), (16, 37, 1)])
    t = env.observe()
    obs = t[0]["obs"]
    assert obs[121] == 46.078141022779725 and obs[122] == 10.0
    assert _grid_nonzeros(obs) == [(3, 6, 1.0), (5, 5, 1.0), (5, 10, 3.0), (6, 0, 2.0)]


def test_kat6b_obs_normal():
    env = scene([(37.5, 37.5, 12)],
                pellets=[(30, 30, 1), (40, 37, 2), (60, 60, 3), (37, 37, 1), (16, 37, 1), (14, 50, 1)])
    t = env.observe()
    obs = t[0]["obs"]
    assert obs[121] == 48.11721642601227 and obs[122] == 12.0
    assert _grid_nonzeros(obs) == [(3, 3, 1.0), (5, 0, 1.0), (5, 5, 3.0), (5, 6, 2.0), (8, 0, 1.0), (10, 10, 3.0)]


def test_kat7_reward_and_frame_skip():
    env = scene([(37.5, 37.5, 12)])
    act = np.array([[0.5, 0.5, 0, 0]], dtype=np.float32)
    acc = []
    for f in range(9):
        t = env.observe()
        if f == 0:
            assert t[0]["need_action"] and not t[0]["valid"]
        elif f < 8:
            assert not t[0]["need_action"] and not t[0]["valid"]
            acc.append(float(env.record.players["bot"]["cum_reward"][0]))
        else:
            assert t[0]["need_action"] and t[0]["valid"] and not t[0]["done"]
            assert t[0]["reward"] == -0.2877761119626392
        env.step(act, 1)
        if f == 0:
            p = env.record.players
            assert (p["cmd_x"][0], p["cmd_y"][0]) == (37.0, 37.0)
    assert acc[0] == pytest.approx(-0.008, abs=1e-12) and acc[-1] == -0.22385072887227153
    assert float(env.record.players["bot"]["last_mass"][0]) == 11.968037308454816


@pytest.mark.parametrize("portable", [False, True])
def test_kats_hold_for_both_math_builds_within_ulps(portable):
    env = scene([(10, 10, 10)], cmd=(20, 10), portable=portable)
    env.field_update()
    assert env.record.cells[0, 0]["x"] == pytest.approx(11.340207150895663, rel=1e-14)
