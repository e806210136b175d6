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
st = b.state_tensor().cpu().numpy()
L = b.layout
pl = np.stack([lay.Record(L, st[e].copy()).players["n_cells"].copy() for e in range(min(E, 4096))])
tot = pl.sum(axis=1)
w = work[:len(tot)]
print("  corr(work, live cells of all players) = %.2f" % np.corrcoef(w, tot)[0, 1])
for lo, hi in ((0, 3), (3, 6), (6, 12), (12, 20), (20, 40), (40, 400)):
    m = (tot >= lo) & (tot < hi)
    if m.any():
        print("  envs with %3d..%3d live cells: %5.1f %% of envs, mean work %8.0f cycles, by phase: %s" % (
            lo, hi - 1, 100.0 * m.mean(), w[m].mean(), " ".join("%s=%.0f" % (names[i].split()[0], c[:len(tot)][m, i].mean()) for i in (3, 4, 5, 6, 7, 9, 10))))
