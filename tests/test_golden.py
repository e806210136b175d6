"""The C oracle replays the golden fixtures that tools/gen_golden.py captured by EXECUTING the reference
(tests/golden/*.npz): records bit for bit, event hash every frame, every observation (float32)."""
import ast
import glob
import os

import numpy as np
import pytest

import aigar_b200.layout as lay
from oracle import oracle as orc

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def replay(path, make_env, rtol=0.0, obs_rtol=0.0, obs_stats=None):
    z = np.load(path)
    kw = ast.literal_eval(str(z["kw"]))
    cfg = lay.derive_config(event_cap=0, **kw)
    env = make_env(cfg, int(z["seed"]), int(z["env_id"]))
    L = env.layout
    rec_at = {int(f): i for i, f in enumerate(z["record_frames"])}
    obs_at = {(int(t), int(a)): i for i, (t, a) in enumerate(z["obs_index"])}
    flags = {(int(f[0]), int(f[1])): tuple(int(v) for v in f[2:]) for f in z["flags"]}
    d = lay.compare_records(lay.Record(L, z["records"][rec_at[-1]].copy()), env.record, rtol=rtol, what="init ")
    assert not d, d
    for t in range(z["actions"].shape[0]):
        turn = env.frame(z["actions"][t])
        for a in range(L.n_agents):
            assert (int(turn[a]["observed"]), int(turn[a]["valid"]), int(turn[a]["done"]), int(turn[a]["need_action"])) \
                == flags[(t, a)], (t, a)
            if (t, a) in obs_at:
                ref = z["obs"][obs_at[(t, a)]]
                got = turn[a]["obs32"]
                assert got is not None
                if obs_stats is not None:
                    obs_stats["n"] += ref.size
                    obs_stats["bad"] += int((~np.isclose(ref, got, rtol=obs_rtol, atol=obs_rtol)).sum())
                elif obs_rtol == 0.0:
                    assert np.array_equal(ref, got), (t, a, np.argwhere(ref != got)[:4])
                else:
                    np.testing.assert_allclose(got, ref, rtol=obs_rtol, atol=obs_rtol)
        assert int(env.record.header["event_hash"][0]) == int(z["event_hash"][t]), "event hash differs at frame %d" % t
        if t in rec_at:
            d = lay.compare_records(lay.Record(L, z["records"][rec_at[t]].copy()), env.record, rtol=rtol,
                                    what="frame %d " % t, check_hist=obs_stats is None)  # history grids are observations
            assert not d, d


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_the_reference_bit_for_bit(path):
    replay(path, lambda cfg, seed, env_id: orc.OracleEnv(cfg, seed=seed, env_id=env_id))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_portable_math_oracle_same_events_state_within_1e9(path):
    """Same eaten / merged / collided sets and pellet indices (event hash), state within 1e-9 relative: the
    distance between libm and include/agar_math.h (the arithmetic the CUDA kernels use).

    Observations: the reference bins objects with int(x / gsSize) evaluated exactly AT bucket edges
    (spatialHashTable.py:91-112), so a last-bit difference in fov / position moves an object by one grid square
    (DESIGN.md "observation conditioning").  Bounded here: < 3 % of observation elements may differ."""
    stats = {"n": 0, "bad": 0}
    replay(path, lambda cfg, seed, env_id: orc.OracleEnv(cfg, seed=seed, env_id=env_id, portable=True), rtol=1e-9,
           obs_rtol=1e-5, obs_stats=stats)
    if "canonical" in os.path.basename(path):  # AGAR_OBS_CANONICAL: exact floors -> no edge flips at all
        assert stats["n"] > 0 and stats["bad"] == 0, stats
    else:
        assert stats["n"] > 0 and stats["bad"] <= 0.03 * stats["n"], stats


def test_fixtures_exist():
    assert len(GOLDEN) >= 4
