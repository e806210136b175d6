"""Find the first decision at which a GPU random-action rollout leaves the portable oracle (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aigar_b200.layout as lay
from aigar_b200.env import AgarBatch
from oracle import oracle as orc

def main(n=48, W=8, decisions=125, chunk=1, seed=77, first=1000):
    cfg = lay.derive_config(event_cap=64)
    b = AgarBatch(cfg, n, seed=seed, first_env_id=first, tile_width=W)
    oras = [orc.OracleEnv(cfg, seed=seed, env_id=first + i, portable=True) for i in range(n)]
    for d in range(0, decisions, chunk):
        b.rollout_random(chunk, 8, d)
        st = b.state_tensor().cpu().numpy()
        for i, e in enumerate(oras):
            e.rollout_random(chunk, 8, d)
            diff = lay.compare_records(e.record, lay.Record(b.layout, st[i].copy()), what="dec %d env %d " % (d, i), check_events=True)
            if diff:
                print("\n".join(diff[:10]))
                print("oracle events", e.record.event_list()); print("gpu events", lay.Record(b.layout, st[i].copy()).event_list())
                c = e.record.cells[0, 0]; print("oracle cell", c["x"], c["y"], c["mass"], c["radius"])
                return False
    print("OK", n, W, decisions, chunk)
    return True

if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    sys.exit(0 if main(*a) else 1)
