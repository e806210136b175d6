"""One short random-action rollout for ncu: python tools/prof_run.py ENVS TILE DECISIONS"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aigar_b200.layout as lay
from aigar_b200.env import AgarBatch
E, W, D = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
b = AgarBatch(lay.derive_config(), E, seed=1, tile_width=W)
for i in range(3):
    b.rollout_random(D, 8, i * D)
torch.cuda.synchronize()
print("ok", E, W, D)
