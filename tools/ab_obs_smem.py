"""A/B of k_simple's observation accumulator: shared-memory staged (default) vs REDs into the caller's row (AGAR_SIMPLE_OBS_SMEM=0).
Run on a GPU box: python tools/ab_obs_smem.py"""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CHILD = r'''
import sys, json, torch
sys.path.insert(0, %r)
import aigar_b200.layout as lay
from aigar_b200.env import AgarBatch
out = {}
for E, W, dec, fpd in ((4096, 8, 125, 8), (4096, 8, 1000, 1), (65536, 2, 25, 8), (1048576, 1, 25, 8), (1048576, 1, 100, 1)):
    b = AgarBatch(lay.derive_config(), E, seed=1, tile_width=W)
    b.rollout_random(dec, fpd, 0)
    torch.cuda.synchronize()
    best = 1e9
    for i in range(4):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); b.rollout_random(dec, fpd, (i + 1) * dec); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    out["%%d envs W=%%d obs every %%d" %% (E, W, fpd)] = E * dec * fpd / (best * 1e-3)
    b.close()
print(json.dumps(out))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

res = {}
for mode in ("1", "0", "1", "0"):
    env = dict(os.environ, AGAR_SIMPLE_OBS_SMEM=mode)
    r = json.loads(subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, check=True).stdout.strip().splitlines()[-1])
    for k, v in r.items():
        res.setdefault(k, {}).setdefault(mode, []).append(v)
for k, v in res.items():
    print("%-34s smem %s   global REDs %s" % (k, ["%.3g" % x for x in v["1"]], ["%.3g" % x for x in v["0"]]))
