"""Drive the reference harness and the C oracle side by side and report the first divergence."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import aigar_b200.layout as lay
from oracle import ref_harness as rh
from oracle import oracle as orc

def run(kw, frames, seed=1, env_id=0, verbose=True, obs_tol=0.0):
    cfg = lay.derive_config(event_cap=512, **kw)
    ref = rh.RefEnv(cfg, seed=seed, env_id=env_id)
    ora = orc.OracleEnv(cfg, seed=seed, env_id=env_id)
    L = ora.layout
    d = lay.compare_records(ref.to_record(), ora.record, what="init ", check_events=True)
    if d:
        print("\n".join(d)); return False
    rng = np.random.default_rng(seed * 1000 + env_id)
    for t in range(frames):
        act = rng.random((max(L.n_agents, 1), 4)).astype(np.float32)
        tr = ref.step(act)
        to = ora.frame(act)
        rrec = ref.to_record(tr)
        d = lay.compare_records(rrec, ora.record, what="frame %d " % t, check_events=True)
        for a in range(L.n_agents):
            for key in ("observed", "valid", "done", "need_action"):
                if tr[a][key] != to[a][key]:
                    d.append("frame %d agent %d %s: %r != %r" % (t, a, key, tr[a][key], to[a][key]))
            if tr[a]["obs"] is not None and to[a]["obs"] is not None:
                if obs_tol == 0.0:  # the API's obs are float32; history channels are stored as float32
                    bad = tr[a]["obs"].astype(np.float32) != to[a]["obs32"]
                else:
                    bad = ~np.isclose(tr[a]["obs"], to[a]["obs"], rtol=obs_tol, atol=obs_tol)
                if bad.any():
                    i = int(np.argwhere(bad)[0][0])
                    d.append("frame %d agent %d obs[%d]: %r != %r (%d bad)" % (t, a, i, tr[a]["obs"][i], to[a]["obs"][i], int(bad.sum())))
            elif (tr[a]["obs"] is None) != (to[a]["obs"] is None):
                d.append("frame %d agent %d obs presence differs" % (t, a))
        if d:
            print("\n".join(d[:20]))
            print("ref events:", ref.events[:20]); print("ora events:", ora.record.event_list()[:20])
            return False
    if verbose:
        print("OK", kw, frames, "frames; event hash %016x" % int(ora.record.header["event_hash"][0]))
    return True

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "1"
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    kws = {"1": dict(), "3": dict(num_nn=1, num_greedy=1, virus=True, split=True, eject=True),
           "4": dict(num_nn=8, num_greedy=8, virus=True, split=True, eject=True),
           "r": dict(num_nn=1, num_greedy=1, num_random=1, virus=True, split=True, eject=True),
           "4nv": dict(num_nn=8, num_greedy=8, virus=False, split=True, eject=True)}
    ok = run(kws[which], frames, seed=seed)
    sys.exit(0 if ok else 1)
