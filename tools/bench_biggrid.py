import sys, json
sys.path.insert(0, '.')
import torch
import aigar_b200.layout as lay
from aigar_b200.env import AgarBatch
for kw, E in ((dict(grid=42, overrides={"use_fovsize": 0, "use_totalmass": 0}), 4096), (dict(grid=42, overrides={"use_fovsize": 0, "use_totalmass": 0}), 65536),
              (dict(grid=42, num_nn=1, num_greedy=1, virus=True, split=True, eject=True), 4096)):
    cfg = lay.derive_config(**kw)
    b = AgarBatch(cfg, E, seed=1)
    b.rollout_random(12, 8, 0)
    torch.cuda.synchronize()
    best = 1e9
    for i in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); b.rollout_random(12, 8, (i + 1) * 12); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    print(json.dumps({"kw": repr(kw), "envs": E, "tile": b.tile_width, "state_len": int(b.layout.state_len), "env_steps_per_s": E * 96 / (best * 1e-3), "ms": best}), flush=True)
    b.close()
