"""Where a DQN tick of examples/train_dqn.py spends its GPU time (eager launches, CUDA events): python tools/tick_breakdown.py [envs] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aigar_b200.layout as lay
from aigar_b200.dqn import make_dqn
from aigar_b200.env import AgarBatch
from aigar_b200.learner import GraphedDQNLoop
from aigar_b200.replay import GpuReplayBuffer

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
cfg = lay.derive_config()
env = AgarBatch(cfg, E, seed=1)
L = env.layout.state_len
net = make_dqn(L, device=env.device, seed=0)
rp = GpuReplayBuffer(1 << 20, L, 1, prioritized=True, alpha=0.6, beta=0.4)
loop = GraphedDQNLoop(env, net, rp, batch_size=B)
env.observe()
for _ in range(12):
    loop._collect()
    loop._learn()
torch.cuda.synchronize()


def timed(name, fn, reps=30):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    print("%-44s %8.1f us" % (name, s.elapsed_time(e) * 1e3 / reps), flush=True)


acts = torch.rand((E, 1, 4), device=env.device)
idx = torch.zeros((E, 1), device=env.device)
timed("step_observe (8 frames + observation)", lambda: env.step_observe(acts, 8))
timed("replay.add_batch (%d transitions)" % E, lambda: rp.add_batch(loop.prev_obs, idx, env.get(lay.GET_REWARD), env.obs, env.get(lay.GET_DONE), env.get(lay.GET_VALID)))
u = torch.rand(B, dtype=torch.float64, device=env.device)
timed("replay.sample (%d)" % B, lambda: rp.sample(u))
s_, a_, r_, s2_, d_, w_, ix = rp.sample(u)
pr = torch.rand(B, dtype=torch.float64, device=env.device) + 1e-4
timed("replay.update_priorities (%d)" % B, lambda: rp.update_priorities(ix, pr))
timed("collect (decide + step + add), eager", loop._collect)
timed("learn (sample + TD step + priorities), eager", loop._learn)
loop.run(20)
torch.cuda.synchronize()
timed("whole tick replayed as one CUDA graph", lambda: loop.run(1), reps=100)
