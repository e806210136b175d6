"""The reference's default run (ALGORITHM = "CACLA", pellet collection) on the B200 env path with E lock-stepped envs:
env -> DLPack obs -> actor (+ Gaussian noise) -> replay (GPU) -> batched CACLA+Var step (a.i.gar_b200/learner.py).

python examples/train_cacla.py [--envs 4096] [--ticks 3000] [--batch 2048]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aigar_b200.layout as lay
from aigar_b200.env import AgarBatch
from aigar_b200.learner import CACLALearner
from aigar_b200.replay import GpuReplayBuffer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--ticks", type=int, default=3000)
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--reset-every", type=int, default=250)
    args = ap.parse_args()
    cfg = lay.derive_config()
    env = AgarBatch(cfg, args.envs, seed=1)
    L = env.layout.state_len
    # NOISE_AT_HALF_TRAINING = 0.02 (networkParameters.py:55-58); learning rates scaled for batches of 2048 instead of 32
    lrn = CACLALearner(L, device=env.device, noise=1.0, noise_decay=0.02 ** (1.0 / (args.ticks / 2)), critic_lr=0.0005,
                       actor_lr=0.0005, max_epochs=4, target_network_steps=200)
    rp = GpuReplayBuffer(1 << 20, L, 2, prioritized=False)
    g = torch.Generator(device=env.device).manual_seed(1)
    acts = torch.zeros((args.envs, 1, 4), device=env.device)
    obs = env.observe().clone()
    t0 = time.time()
    for tick in range(args.ticks):
        _, noisy = lrn.decide(obs.view(args.envs, L))
        acts[:, 0, :2] = noisy
        nxt = env.step_observe(acts, cfg.frame_skip + 1)
        rp.add_batch(obs, noisy, env.get(lay.GET_REWARD), nxt, env.get(lay.GET_DONE), env.get(lay.GET_VALID))
        obs = nxt.clone()
        if tick >= 8:
            s, a, r, s2, d, ix = rp.sample(torch.rand(args.batch, dtype=torch.float64, device=env.device, generator=g))
            td, loss, epochs = lrn.learn(s, a, r, s2, d)
        if (tick + 1) % args.reset_every == 0:
            mass = env.get(lay.GET_MASS).mean().item()
            print("tick %5d  noise %.3f  mean mass after %d frames %.1f  critic loss %.3f  (%.0f env-steps/s incl. learning)" % (
                tick + 1, lrn.std, args.reset_every * 8, mass, loss, args.envs * (tick + 1) * 8 / (time.time() - t0)), flush=True)
            env.reset()
            env.reset_bots()
            obs = env.observe().clone()


if __name__ == "__main__":
    main()
