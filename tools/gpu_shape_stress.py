"""Long steady-state cross-check of the multi-agent kernel's fast paths (run on a GPU box): the 32-lane shape goes through the
per-env pellet index, the cooperative self-collision sweep and the cooperative player-player pass; the 16-lane shape runs the
sequential lane-0 forms and scans the pellet pool directly.  Same seeds, same global env ids -> records must be IDENTICAL after
every chunk; a few envs are also replayed by the CPU oracle.  python tools/gpu_shape_stress.py [frames] [envs]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aigar_b200.layout as lay
from aigar_b200.env import AgarBatch
from oracle import oracle as orc

CASES = [dict(num_nn=8, num_greedy=8, virus=True, split=True, eject=True),
         dict(num_nn=4, num_greedy=3, num_random=1, virus=True, split=True, eject=True),
         dict(num_nn=2, num_greedy=4, virus=False, split=True, eject=True),
         dict(num_nn=1, num_greedy=1, virus=True, split=True, eject=True)]


def run_case(ci, kw, frames, n_envs):
    cfg = lay.derive_config(**kw)
    a = AgarBatch(cfg, n_envs, seed=41 + ci, first_env_id=1000)
    b = AgarBatch(cfg, n_envs, seed=41 + ci, first_env_id=1000, tile_width=16)
    t0 = time.time()
    for d in range(0, frames // 8, 25):
        a.rollout_random(25, 8, d)
        b.rollout_random(25, 8, d)
        diff = (a.state_tensor() != b.state_tensor()).any(dim=1).nonzero().flatten().tolist()
        if diff:
            print("case %d %r: 32-lane and 16-lane shapes differ after frame %d in envs %r" % (ci, kw, (d + 25) * 8, diff[:8]))
            return False
    ok = True
    st = a.state_tensor().cpu().numpy()
    done = (frames // 8 + 24) // 25 * 25
    for e in (0, n_envs // 2, n_envs - 1):
        o = orc.OracleEnv(cfg, seed=41 + ci, env_id=1000 + e, portable=True)
        o.rollout_random(done, 8, 0)
        dd = lay.compare_records(o.record, lay.Record(a.layout, st[e].copy()), what="env %d " % e)
        if dd:
            print("case %d env %d differs from the oracle: %r" % (ci, e, dd[:4]))
            ok = False
    print("case %d %r: %d envs x %d frames identical in both shapes and equal to the oracle (%.0fs, pellet index %s)" % (
        ci, kw, n_envs, done * 8, time.time() - t0, "on" if a.layout.pellet_cap > 256 else "off"), flush=True)
    return ok


def run(frames=2000, n_envs=300, cases=None):
    ok = True
    for ci, kw in enumerate(CASES):
        if cases is None or ci in cases:
            ok = run_case(ci, kw, frames, n_envs) and ok
    return ok


if __name__ == "__main__":
    sys.exit(0 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 2000, int(sys.argv[2]) if len(sys.argv) > 2 else 300) else 1)
