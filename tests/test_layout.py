"""Host logic: record layout (C header == Python twin), struct sizes, the closed-form pellet hash rectangle."""
import ctypes

import numpy as np
import pytest

import aigar_b200.layout as lay
from oracle import oracle as orc

CONFIGS = [dict(), dict(num_nn=1, num_greedy=1, virus=True, split=True, eject=True),
           dict(num_nn=8, num_greedy=8, virus=True, split=True, eject=True),
           dict(num_nn=8, num_greedy=8, virus=False, split=True, eject=True),
           dict(num_nn=2, num_random=1, split=True),
           dict(grid_view=False), dict(num_nn=3, num_greedy=5, virus=True, split=True, eject=True, grid_view=False)]


@pytest.mark.parametrize("kw", CONFIGS)
def test_c_layout_equals_python_layout(kw):
    cfg = lay.derive_config(event_cap=32, **kw)
    c = lay.AgarLayout()
    assert orc.load().oracle_layout(ctypes.byref(cfg), ctypes.byref(c)) == 0
    assert c.as_dict() == lay.layout_for_config(cfg).as_dict()
    assert c.record_bytes % 128 == 0


def test_sizes_match_the_survey():
    for k, (s, p, l) in {1: (75, 85, 123), 2: (106, 169, 854), 16: (300, 1350, 854)}.items():
        cfg = lay.derive_config(num_nn=1, num_greedy=k - 1, virus=k > 1, split=k > 1, eject=k > 1)
        L = lay.layout_for_config(cfg)
        assert (L.field_size, L.pellet_cap, L.state_len) == (s, p, l)
    cfg = lay.derive_config(num_nn=8, num_greedy=8, virus=False, split=True, eject=True)
    assert lay.layout_for_config(cfg).state_len == 733
    L = lay.layout_for_config(lay.derive_config(num_nn=2, num_greedy=2, split=True, eject=True, grid_view=False))
    assert (L.state_len, L.n_grids, L.n_extra, L.n_hist) == (lay.SIMPLE_STATE_LEN, 0, 0, 0)  # bot.py:511-548: 12 values


def test_invalid_configs_are_rejected():
    lib = orc.load()
    out = lay.AgarLayout()
    bad = lay.derive_config()
    bad.n_players = 0
    assert lib.oracle_layout(ctypes.byref(bad), ctypes.byref(out)) == -1
    bad = lay.derive_config(eject=True)  # eject without split raises in the reference (bot.py:568)
    assert lib.oracle_layout(ctypes.byref(bad), ctypes.byref(out)) == -4
    bad = lay.derive_config(num_nn=1, num_greedy=1)
    bad.bot_type[0], bad.bot_type[1] = lay.BOT_GREEDY, lay.BOT_NN  # NN bots must come first (aigar.py:778-780)
    assert lib.oracle_layout(ctypes.byref(bad), ctypes.byref(out)) == -1


def test_pellet_rectangle_closed_form():
    """The kernels use an integer formula for the hash rectangle of an integer pellet; prove it equal to
    spatialHashTable.getIdsForArea (float arithmetic) for every coordinate, mass and field size."""
    lib = orc.load()
    b0, b1 = ctypes.c_int(), ctypes.c_int()
    for S in (75, 106, 300, 1023):
        for m in (1, 2, 3):
            r = float(np.sqrt(m / np.pi))
            for x in range(S):
                lib.oracle_axis_range(float(x), r, S, ctypes.byref(b0), ctypes.byref(b1))
                assert (b0.value, b1.value) == ((x - 1 if x > 0 else 0) // 20, x // 20), (S, m, x)


def test_eat_test_implies_shared_bucket():
    """k_simple skips the hash-rectangle test for the first pellet a cell eats in a frame: whenever the eat test passes
    (cell.py:143-152 overlap with the cell as the bigger one) the pellet's rectangle meets the cell's (pre-growth) one; and it
    passes the kernel's float32 candidate filter."""
    lib = orc.load()
    rng = np.random.default_rng(5)
    a0, a1, c0, c1 = (ctypes.c_int() for _ in range(4))
    checked = 0
    for S in (75, 106, 300):
        for _ in range(6000):
            mass = float(rng.choice([rng.uniform(4, 400), rng.uniform(400, 22500)]))
            r = float(np.sqrt(mass / np.pi))
            x, y = rng.uniform(0, S, 2)
            if rng.random() < 0.3:  # hug a wall or a bucket edge
                x = float(rng.choice([0.0, S, 20.0 * rng.integers(0, S // 20 + 1)])) + rng.uniform(-1e-9, 1e-9)
                x = min(max(x, 0.0), float(S))
            for _ in range(8):
                ang, d = rng.uniform(0, 2 * np.pi), r / np.sqrt(1.1) * rng.choice([rng.uniform(0, 1), rng.uniform(0.999, 1.001)])
                px, py = int(round(x + d * np.cos(ang))), int(round(y + d * np.sin(ang)))
                if not (0 <= px < S and 0 <= py < S):
                    continue
                d2 = (x - px) * (x - px) + (y - py) * (y - py)
                if not d2 * 1.1 < r * r:
                    continue
                checked += 1
                lib.oracle_axis_range(float(x), r, S, ctypes.byref(a0), ctypes.byref(a1))
                lib.oracle_axis_range(float(y), r, S, ctypes.byref(c0), ctypes.byref(c1))
                assert a0.value <= px // 20 <= a1.value and c0.value <= py // 20 <= c1.value, (S, mass, x, y, px, py)
                # ... and it passes both of k_simple's candidate filters (agar_simple.cuh, s_field_update): the integer window
                assert abs(px - int(x)) <= int(r) + 1 and abs(py - int(y)) <= int(r) + 1, (mass, x, y, px, py)
                # ... and the float32 disc
                f = np.float32
                dxf, dyf = f(px) - f(x), f(py) - f(y)
                assert f(dxf * dxf) + f(dyf * dyf) <= f(0.9101) * f(r * r) + f(0.1), (mass, x, y, px, py)
    assert checked > 20000


def test_sharding_partition():
    from aigar_b200.sharding import shard_envs
    for total in (1, 7, 4096, 65536 + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_envs(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == total
            for (f0, n0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + n0 == f1
