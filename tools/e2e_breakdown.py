"""Where the time of one host-buffer step goes (run on a GPU box): python tools/e2e_breakdown.py [envs]
 a) device buffers only: agar_step_observe + stream sync                      (kernel + launch + sync)
 b) agar_step_host, pinned buffers, reward / done only (obs_host = NULL)      (+ flag polling instead of the sync)
 c) agar_step_host, pinned buffers, observations too                          (+ E x L x 4 bytes over PCIe from the CTAs)
 d) c) in G env groups on their own streams, launches paced                   (PCIe of one group under the frames of the others)
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aigar_b200.layout as lay
from aigar_b200.env import AgarBatch

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = lay.derive_config()
N = 400


def timed(fn, n=N):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


b = AgarBatch(cfg, E, seed=1)
L = b.layout
acts_d = torch.rand((E, 1, 4), device="cuda")
b.observe()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
b.step_observe(acts_d, 8)
torch.cuda.synchronize()
s.record()
for _ in range(N):
    b.step_observe(acts_d, 8)
e.record()
torch.cuda.synchronize()
print("kernel alone (back-to-back launches, CUDA events): %.1f us per 8-frame step of %d envs" % (s.elapsed_time(e) / N * 1e3, E))


def dev():
    b.step_observe(acts_d, 8)
    torch.cuda.synchronize()


print("a) device buffers + sync:          %.1f us" % timed(dev))
acts = torch.rand((E, 1, 4)).pin_memory()
obs_h = torch.empty((E, 1, L.state_len)).pin_memory()
rew_h = torch.empty((E, 1)).pin_memory()
done_h = torch.empty((E, 1), dtype=torch.uint8).pin_memory()
ap, op, rp, dp = acts.data_ptr(), obs_h.data_ptr(), rew_h.data_ptr(), done_h.data_ptr()


def host_noobs():
    b.step_host_begin_ptr(ap, 8, None)
    b.step_host_end_ptr(rp, dp)


def host_obs():
    b.step_host_begin_ptr(ap, 8, op)
    b.step_host_end_ptr(rp, dp)


print("b) pinned, reward/done only, poll: %.1f us" % timed(host_noobs))
print("c) pinned, + observations (%.2f MB): %.1f us" % (obs_h.numel() * 4 / 1e6, timed(host_obs)))
b.close()
for G in (2, 4):
    Eg = E // G
    hs = [AgarBatch(cfg, Eg, seed=1, first_env_id=g * Eg, stream=torch.cuda.Stream()) for g in range(G)]
    for h in hs:
        h.observe()
    ptrs = [(acts[g * Eg:].data_ptr(), obs_h[g * Eg:].data_ptr(), rew_h[g * Eg:].data_ptr(), done_h[g * Eg:].data_ptr()) for g in range(G)]
    for spacing_us in (0, 10, 20, 30, 40):
        sp = spacing_us * 1e-6
        last = [0.0]

        def go(g):
            while time.perf_counter() - last[0] < sp:
                pass
            hs[g].step_host_begin_ptr(ptrs[g][0], 8, ptrs[g][1])
            last[0] = time.perf_counter()

        for g in range(G):
            go(g)
        torch.cuda.synchronize
        t0 = time.perf_counter()
        for _ in range(N):
            for g in range(G):
                hs[g].step_host_end_ptr(ptrs[g][2], ptrs[g][3])
                go(g)
        for g in range(G):
            hs[g].step_host_end_ptr(ptrs[g][2], ptrs[g][3])
        dt = (time.perf_counter() - t0) / N * 1e6
        print("d) %d groups, spacing %2d us: %.1f us per step of all %d envs" % (G, spacing_us, dt, E))
    for h in hs:
        h.close()
