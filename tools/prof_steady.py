"""Steady-state launch of a multi-agent config for ncu: one warm-up rollout (cells grown, splits / merges / viruses live), then
the launch to profile.  python tools/prof_steady.py CONFIG ENVS TILE WARM_DECISIONS DECISIONS
ncu -k regex:k_main --launch-skip 1 --launch-count 1 ... picks the second k_main launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aigar_b200.layout as lay
from aigar_b200.env import AgarBatch
from sweep import KWS
which, E, W, D0, D = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
b = AgarBatch(lay.derive_config(**KWS[which]), E, seed=1, tile_width=W)
b.rollout_random(D0, 8, 0)
torch.cuda.synchronize()
b.rollout_random(D, 8, D0)
torch.cuda.synchronize()
print("ok", which, E, W, D0, D)
