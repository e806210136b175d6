"""Per-phase cycle breakdown of the multi-agent kernel at steady state (debug build with -DAGAR_PHASE_CLOCKS).  The frame
barrier's wait shows up in the slot AFTER the barrier (verified by removing the fov pass: the ~1 M cycles moved to the next slot):
   nvcc ... -DAGAR_PHASE_CLOCKS -o ab/libagar_clk.so ; AGAR_B200_LIB=ab/libagar_clk.so python tools/phase_clocks.py CONFIG ENVS"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aigar_b200.layout as lay
from aigar_b200.env import AgarBatch, load_library
from sweep import KWS
which, E = sys.argv[1], int(sys.argv[2])
FS = int(sys.argv[3]) if len(sys.argv) > 3 else 7
b = AgarBatch(lay.derive_config(frame_skip=FS, **KWS[which]), E, seed=2026, first_env_id=3 * 10 ** 6)
lib = load_library()
b.rollout_random(600 // (FS + 1), FS + 1, 0)
torch.cuda.synchronize()
buf = np.zeros((E, 16), dtype=np.uint64)
lib.agar_debug_read_clocks(ctypes.c_void_p(buf.ctypes.data), E, 1)
b.rollout_random(200 // (FS + 1), FS + 1, 75)
torch.cuda.synchronize()
lib.agar_debug_read_clocks(ctypes.c_void_p(buf.ctypes.data), E, 1)
names = ["(clock read hoisted above the barrier)", "NN turn end (cmd point)", "wait field barrier", "ph0 viruses/blobs/players", "ph1 merge/virus ovl", "ph2 pellets",
         "ph3 blob/player-player/spawn", "scripted turns", "WAIT AT THE FRAME BARRIER + fov pass", "NN bookkeeping", "NN observe", "live list", "frame tail"]
c = buf.astype(np.float64) / float(200 // (FS + 1) * (FS + 1))  # cycles per frame
tot = c.sum(axis=1)
print("config %s, %d envs, frames 600-800: mean cycles per frame per tile %.0f" % (which, E, tot.mean()))
for i, n in enumerate(names):
    print("  %-30s mean %8.0f (%4.1f%%)   p95 %8.0f   max %8.0f" % (n, c[:, i].mean(), 100 * c[:, i].mean() / tot.mean(), np.percentile(c[:, i], 95), c[:, i].max()))
work = c[:, [1, 3, 4, 5, 6, 7, 9, 10, 11, 12]].sum(axis=1)
print("  work (no waits): mean %.0f  p50 %.0f  p95 %.0f  max %.0f" % (work.mean(), np.median(work), np.percentile(work, 95), work.max()))
ncells = b.get(lay.GET_NCELLS).cpu().numpy().sum(axis=1) if b.layout.n_agents else None
if ncells is not None:
    print("  corr(work, agents' cell count) = %.2f" % np.corrcoef(work, ncells)[0, 1])
