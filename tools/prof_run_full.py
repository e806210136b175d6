"""Short random-action rollout of a multi-agent config for ncu: python tools/prof_run_full.py CONFIG ENVS TILE DECISIONS"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aigar_b200.layout as lay
from aigar_b200.env import AgarBatch
from sweep import KWS
which, E, W, D = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
b = AgarBatch(lay.derive_config(**KWS[which]), E, seed=1, tile_width=W)
for i in range(3):
    b.rollout_random(D, 8, i * D)
torch.cuda.synchronize()
print("ok", which, E, W, D)
