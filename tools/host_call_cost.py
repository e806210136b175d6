"""Host-side cost of the host-buffer calls (the launch is asynchronous: the wall time of agar_step_host_begin is what the calling
thread pays per group and tick).  python tools/host_call_cost.py [envs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aigar_b200.layout as lay
from aigar_b200.env import AgarBatch

E = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cfg = lay.derive_config()
st = torch.cuda.Stream()
b = AgarBatch(cfg, E, seed=1, stream=st)
L = b.layout
acts = torch.rand((64, E, 1, 4)).pin_memory()
obs = torch.empty((E, 1, L.state_len)).pin_memory()
rew = torch.empty((E, 1)).pin_memory()
done = torch.empty((E, 1), dtype=torch.uint8).pin_memory()
a_ptr = [acts[i].data_ptr() for i in range(64)]
o, r, d = obs.data_ptr(), rew.data_ptr(), done.data_ptr()
b.observe()
for i in range(20):
    b.step_host_begin_ptr(a_ptr[i % 64], 8, o)
    b.step_host_end_ptr(r, d)
tb = te = tw = 0.0
N = 2000
clock = time.perf_counter
for i in range(N):
    t0 = clock()
    b.step_host_begin_ptr(a_ptr[i % 64], 8, o)
    t1 = clock()
    torch.cuda.synchronize()       # the step is complete: what remains of _end is pure host work
    t2 = clock()
    b.step_host_end_ptr(r, d)
    t3 = clock()
    tb += t1 - t0
    tw += t2 - t1
    te += t3 - t2
print("agar_step_host_begin: %.2f us of host time per call; kernel + export (sync wait) %.1f us; agar_step_host_end after completion: %.2f us" % (
    tb / N * 1e6, tw / N * 1e6, te / N * 1e6))
lib, h = b.lib, b.h
s = st.cuda_stream
t0 = clock()
for i in range(N):
    lib.agar_get_tile_width(h)
print("a trivial ctypes call: %.2f us" % ((clock() - t0) / N * 1e6))
