"""Where does a synchronous agar_step_host call spend its time?  (4096 envs, 8 frames per call)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aigar_b200.layout as lay
from aigar_b200.env import AgarBatch

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = lay.derive_config()
b = AgarBatch(cfg, E, seed=1)
L = b.layout
acts = torch.rand((64, E, 1, 4)).pin_memory()
acts_np = acts.numpy()
obs_h = torch.empty((E, 1, L.state_len)).pin_memory()
rew_h = torch.empty((E, 1)).pin_memory()
done_h = torch.empty((E, 1), dtype=torch.uint8).pin_memory()
d_act = torch.rand((E, 1, 4), device=b.device)
N = 400


def timeit(name, f):
    for i in range(20):
        f(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(N):
        f(i)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / N
    print("%-46s %8.1f us/call  -> %.3g env-steps/s" % (name, dt * 1e6, E * 8 / dt), flush=True)


o, r, d = obs_h.numpy(), rew_h.numpy(), done_h.numpy()
timeit("step_host (actions H2D, obs+reward+done D2H)", lambda i: b.step_host(acts_np[i % 64], 8, o, r, d))
timeit("step_observe on device + sync", lambda i: (b.step_observe(d_act, 8), torch.cuda.synchronize()))
timeit("step_observe on device, no sync", lambda i: b.step_observe(d_act, 8))
dob = b.observe()
timeit("obs D2H only + sync", lambda i: (obs_h.copy_(dob.view_as(obs_h), non_blocking=True), torch.cuda.synchronize()))
timeit("actions H2D only + sync", lambda i: (d_act.copy_(acts[i % 64], non_blocking=True), torch.cuda.synchronize()))
for W in (4, 16, 32):
    try:
        b.set_tile_width(W)
        timeit("step_host, tile width %d" % W, lambda i: b.step_host(acts_np[i % 64], 8, o, r, d))
    except Exception as ex:
        print("W", W, ex)

# ---- G groups of E/G envs, each handle on its own stream, pipelined through agar_step_host_begin / _end
for G in (2, 4):
    n = E // G
    groups = [AgarBatch(cfg, n, seed=1, first_env_id=g * n) for g in range(G)]
    streams = [torch.cuda.Stream() for _ in range(G)]
    bufs = [(torch.empty((n, 1, L.state_len)).pin_memory().numpy(), torch.empty((n, 1)).pin_memory().numpy(),
             torch.empty((n, 1), dtype=torch.uint8).pin_memory().numpy()) for _ in range(G)]

    def begin(g, i):
        with torch.cuda.stream(streams[g]):
            groups[g].step_host_begin(acts_np[i % 64, g * n:(g + 1) * n], 8, bufs[g][0])

    def end(g):
        with torch.cuda.stream(streams[g]):
            groups[g].step_host_end(bufs[g][1], bufs[g][2])

    for g in range(G):
        begin(g, 0)

    def cycle(i):
        for g in range(G):
            end(g)          # group g's observations are on the host: a policy would produce its next actions here
            begin(g, i + 1)

    timeit("%d groups of %d envs pipelined (begin/end)" % (G, n), cycle)
    for g in range(G):
        end(g)

# ---- CPU cost of the enqueue half against the wait half (one group)
tb = te = 0.0
for i in range(N):
    t0 = time.perf_counter()
    b.step_host_begin(acts_np[i % 64], 8, o)
    t1 = time.perf_counter()
    b.step_host_end(r, d)
    t2 = time.perf_counter()
    tb += t1 - t0
    te += t2 - t1
print("one group: begin (enqueue, CPU) %.1f us, end (wait + unpack) %.1f us" % (tb / N * 1e6, te / N * 1e6))
lib, h, st = b.lib, b.h, b._stream()
pa, po, pr, pd = acts_np[0].ctypes.data, o.ctypes.data, r.ctypes.data, d.ctypes.data
t0 = time.perf_counter()
for i in range(N):
    lib.agar_step_host(h, pa, 8, po, pr, pd, st)
print("raw ctypes agar_step_host, pointers precomputed: %.1f us/call" % ((time.perf_counter() - t0) / N * 1e6))
