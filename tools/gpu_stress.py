"""Long GPU-vs-oracle rollouts of the multi-agent configs with an event-type tally (run on a GPU box)."""
import collections, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aigar_b200.layout as lay
from aigar_b200.env import AgarBatch
from oracle import oracle as orc
from gpu_check import KWS

def run(which, n_envs, decisions, seed=5, tile=None, chunk=5):
    cfg = lay.derive_config(event_cap=512, **KWS[which])
    b = AgarBatch(cfg, n_envs, seed=seed, first_env_id=77, tile_width=tile)
    oras = [orc.OracleEnv(cfg, seed=seed, env_id=77 + i, portable=True) for i in range(n_envs)]
    tally = collections.Counter()
    t0 = time.time()
    for d in range(0, decisions, chunk):
        b.rollout_random(chunk, 8, d)
        st = b.state_tensor().cpu().numpy()
        for i, e in enumerate(oras):
            for dd in range(chunk):
                e.rollout_random(1, 8, d + dd)
            for ev in e.record.event_list():
                tally[lay.EV_NAMES[ev[0]]] += 1
            diff = lay.compare_records(e.record, lay.Record(b.layout, st[i].copy()), what="%s dec %d env %d " % (which, d, i))
            if diff:
                print("\n".join(diff[:8])); return False
    print("OK config %s: %d envs x %d frames bit-exact (%.0fs); events in sampled frames: %s" % (
        which, n_envs, decisions * 8, time.time() - t0, dict(tally)))
    return True

if __name__ == "__main__":
    ok = run("3", 48, 250) and run("r", 24, 200) and run("4", 6, 75) and run("3c", 16, 120) and run("1", 64, 400, tile=8) and run("1", 64, 250, tile=2)
    sys.exit(0 if ok else 1)
