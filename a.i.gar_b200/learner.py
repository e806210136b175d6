"""Batched Q-learning step on replay samples — the consumer of the env path, not part of it.

Replaces the per-sample Python loop of src/model/qLearning.py:146-185 (`train`), `calculateTarget` :122-127 and
`calculateTargetForAction` :114-120 with batched torch ops (library GEMMs): hyper-parameters from
src/model/networkParameters.py (DISCOUNT 0.9, ALPHA 0.001, Adam, TARGET_NETWORK_STEPS 1500, MSE with importance weights)."""
import copy


class DQNLearner(object):
    def __init__(self, net, discount=0.90, lr=0.001, target_network_steps=1500):
        import torch
        self.torch = torch
        self.net = net.train()
        self.target = copy.deepcopy(net).eval()
        for p in self.target.parameters():
            p.requires_grad_(False)
        self.discount = float(discount)
        self.opt = torch.optim.Adam(self.net.parameters(), lr=lr)
        self.target_network_steps = int(target_network_steps)
        self.step = 0

    def targets_and_td(self, obs_t, action_idx, reward, obs_tp1, done):
        """calculateTarget for a whole batch: (targets [B, n_actions], td_error [B]).
        alive = the new state is not None = done == 0 (src/model/qLearning.py:165-172)."""
        torch = self.torch
        with torch.no_grad():
            q_old = self.net(obs_t)
            q_next = self.target(obs_tp1)
            alive = (done == 0).to(q_old.dtype)
            updated = reward.to(q_old.dtype) + self.discount * q_next.max(dim=1).values * alive
            a = action_idx.long().view(-1, 1)
            td = updated - q_old.gather(1, a).squeeze(1)
            targets = q_old.clone()
            targets.scatter_(1, a, updated.view(-1, 1))
        return targets, td

    def learn(self, obs_t, action_idx, reward, obs_tp1, done, weights=None):
        """QLearn.learn (:187-193): one Adam step on MSE(pred, targets) with optional importance weights; returns the TD
        errors (the new priorities are |td| + 1e-4, src/aigar.py:1081)."""
        torch = self.torch
        targets, td = self.targets_and_td(obs_t, action_idx, reward, obs_tp1, done)
        pred = self.net(obs_t)
        per_sample = ((pred - targets) ** 2).mean(dim=1)  # Keras 'mse': mean over outputs, then sample weights
        loss = (per_sample * weights.to(per_sample.dtype)).mean() if weights is not None else per_sample.mean()
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        self.step += 1
        if self.step % self.target_network_steps == 0:
            self.target.load_state_dict(self.net.state_dict())
        return td, float(loss.detach())
