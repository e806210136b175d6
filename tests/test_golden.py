"""The C oracle replays the golden fixtures that tools/gen_golden.py captured by EXECUTING the reference
(tests/golden/*.npz): records bit for bit, event hash every frame, every observation (float32)."""
import os

import numpy as np
import pytest

import aigar_b200.layout as lay
from oracle import oracle as orc

from golden_util import GOLDEN, LONG, Fixture, assert_event_floor, tally_events


def replay(path, make_env, rtol=0.0, obs_rtol=0.0, obs_stats=None, tally=None):
    fx = Fixture(path)
    cfg = fx.config(event_cap=1024 if tally is not None else 0)
    env = make_env(cfg, fx.seed, fx.env_id)
    L = env.layout
    d = lay.compare_records(fx.record(-1), env.record, rtol=rtol, what="init ")
    assert not d, d
    for t in range(fx.frames):
        turn = env.frame(fx.actions[t])
        if tally is not None:
            tally_events(env.record, tally)
        for a in range(L.n_agents):
            assert (int(turn[a]["observed"]), int(turn[a]["valid"]), int(turn[a]["done"]), int(turn[a]["need_action"])) \
                == fx.flags(t, a), (t, a)
            if (t, a) in fx.obs_at:
                ref = fx.obs[fx.obs_at[(t, a)]]
                got = turn[a]["obs32"]
                assert got is not None
                if obs_stats is not None:
                    obs_stats["n"] += ref.size
                    obs_stats["bad"] += int((~np.isclose(ref, got, rtol=obs_rtol, atol=obs_rtol)).sum())
                elif obs_rtol == 0.0:
                    assert np.array_equal(ref, got), (t, a, np.argwhere(ref != got)[:4])
                else:
                    np.testing.assert_allclose(got, ref, rtol=obs_rtol, atol=obs_rtol)
        assert int(env.record.header["event_hash"][0]) == int(fx.event_hash[t]), "event hash differs at frame %d" % t
        if t in fx.rec_at:
            d = lay.compare_records(fx.record(t), env.record, rtol=rtol, what="frame %d " % t,
                                    check_hist=obs_stats is None)  # history grids are observations
            assert not d, d
    return fx


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_the_reference_bit_for_bit(path):
    replay(path, lambda cfg, seed, env_id: orc.OracleEnv(cfg, seed=seed, env_id=env_id))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_portable_math_oracle_reproduces_the_reference_bit_for_bit(path):
    """include/agar_math.h — the arithmetic of the CUDA kernels — against the reference's own run: every record field, the
    event hash every frame and every float32 observation IDENTICAL to the unpatched reference's, over the 4000-frame 1v1 and
    1600-frame arena rollouts as well.  Round 1's portable math was a few ULP from libm: observations then differed in up to
    3 % of elements (the 12-column defect of spatialHashTable.py:19 hangs on the last bit of fov) and the self-collision
    push-apart amplified the direction's last bit until the event hash left the reference's at frame ~2400.  agar_pow /
    agar_atan2 / agar_sin / agar_cos now restate glibc's algorithms bit for bit (tests/test_portable_math.py)."""
    replay(path, lambda cfg, seed, env_id: orc.OracleEnv(cfg, seed=seed, env_id=env_id, portable=True))


def test_long_fixtures_exercise_every_event_type():
    """The reference-generated long rollouts (1v1: 4000 frames, arena: 1600, pellet config: 1200) replayed through the C
    oracle with its event log on: the oracle's tally equals the tally of the reference's own event recorder, and every
    event type of include/agar_b200.h — split, eject, merge, virus explosion, blob eating, virus-eats-blob, blob -> pellet,
    cell eating, death, respawn — occurs at least MIN_EVENTS_EACH times."""
    total = {}
    assert len(LONG) >= 3
    for path in LONG:
        tally = {}
        fx = replay(path, lambda cfg, seed, env_id: orc.OracleEnv(cfg, seed=seed, env_id=env_id), tally=tally)
        assert {k: v for k, v in tally.items() if v} == {k: v for k, v in fx.tally.items() if v}, (fx.name, tally, fx.tally)
        for k, v in tally.items():
            total[k] = total.get(k, 0) + v
    print("events in the long reference fixtures:", total)
    assert_event_floor(total)


def test_fixtures_exist():
    assert len(GOLDEN) >= 4
