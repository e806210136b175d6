"""End-to-end DQN on the B200 env path: env -> DLPack obs -> torch DQN -> replay (GPU, prioritized) -> batched TD step.
The reference's default run (Q-learning, pellet collection, src/model/networkParameters.py) with E lock-stepped envs.

python examples/train_dqn.py [--envs 4096] [--ticks 3000] [--batch 2048]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aigar_b200.layout as lay
from aigar_b200.dqn import DQNDriver, make_dqn
from aigar_b200.env import AgarBatch
from aigar_b200.learner import DQNLearner
from aigar_b200.replay import GpuReplayBuffer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--ticks", type=int, default=3000)
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--reset-every", type=int, default=250)  # RESET_LIMIT / 8 in the reference is 2500 ticks; shorter here
    ap.add_argument("--no-graph", action="store_true", help="the round-1 loop: one launch per op, a host sync per tick")
    args = ap.parse_args()
    if not args.no_graph:
        return main_graphed(args)
    cfg = lay.derive_config()
    env = AgarBatch(cfg, args.envs, seed=1)
    L = env.layout.state_len
    net = make_dqn(L, device=env.device, seed=0)
    drv = DQNDriver(env, net=net, epsilon=1.0, seed=0)
    learner = DQNLearner(net)
    rp = GpuReplayBuffer(1 << 20, L, 1, prioritized=True, alpha=0.6, beta=0.4)
    g = torch.Generator(device=env.device).manual_seed(1)
    obs = env.observe().clone()
    t0 = time.time()
    for tick in range(args.ticks):
        drv.epsilon = max(0.05, 1.0 - tick / (0.6 * args.ticks))
        net.eval()
        acts = drv.decide()
        idx = drv.last_idx.float().unsqueeze(-1)
        nxt = env.step_observe(acts, cfg.frame_skip + 1)
        rp.add_batch(obs, idx, env.get(lay.GET_REWARD), nxt, env.get(lay.GET_DONE), env.get(lay.GET_VALID))
        obs = nxt.clone()
        if tick >= 8:
            net.train()
            s, a, r, s2, d, w, ix = rp.sample(torch.rand(args.batch, dtype=torch.float64, device=env.device, generator=g))
            td, loss = learner.learn(s, a.squeeze(1), r, s2, d, w.float())
            rp.update_priorities(ix, td.abs().double() + 1e-4)
        if (tick + 1) % args.reset_every == 0:
            mass = env.get(lay.GET_MASS).mean().item()
            print("tick %5d  eps %.2f  mean mass after %d frames %.1f  loss %.3f  (%.0f env-steps/s incl. learning)" % (
                tick + 1, drv.epsilon, args.reset_every * 8, mass, loss, args.envs * (tick + 1) * 8 / (time.time() - t0)), flush=True)
            env.reset()
            env.reset_bots()
            obs = env.observe().clone()


def main_graphed(args):
    """The same run with the whole tick (decide -> step -> replay add -> PER sample -> TD step -> priorities) replayed as one
    CUDA graph (aigar_b200.learner.GraphedDQNLoop): no launch overhead, no host synchronisation inside a tick."""
    from aigar_b200.learner import GraphedDQNLoop
    cfg = lay.derive_config()
    env = AgarBatch(cfg, args.envs, seed=1)
    L = env.layout.state_len
    net = make_dqn(L, device=env.device, seed=0)
    rp = GpuReplayBuffer(1 << 20, L, 1, prioritized=True, alpha=0.6, beta=0.4)
    loop = GraphedDQNLoop(env, net, rp, batch_size=args.batch, eps_decay_ticks=0.6 * args.ticks)
    env.observe()
    torch.cuda.synchronize()
    t0 = t_prev = time.time()
    done = 0
    while done < args.ticks:
        n = min(args.reset_every, args.ticks - done)
        n = loop.run(n)
        done += n
        mass = env.get(lay.GET_MASS).mean().item()  # the only host synchronisation: once per reporting interval
        eps = max(0.05, 1.0 - done / (0.6 * args.ticks) * 0.95)
        now = time.time()
        print("tick %5d  eps %.2f  mean mass after %d frames %.1f  loss %.3f  (%.3g env-steps/s incl. learning in this interval, "
              "%.3g since start incl. graph capture)" % (done, eps, n * 8, mass, float(loop.loss), args.envs * n * 8 / (now - t_prev),
                                                         args.envs * done * 8 / (now - t0)), flush=True)
        t_prev = time.time()
        env.reset()
        env.reset_bots()
        env.observe()
    rp.raise_on_error()


if __name__ == "__main__":
    main()
