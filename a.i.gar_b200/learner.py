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


def make_mlp(in_len, layers, out_len, out_act=None, device="cuda", seed=0):
    """MLP of src/model/network.py (relu hidden, glorot_uniform kernels, zero biases) with `out_act` in (None, "sigmoid")."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    mods, prev = [], in_len
    for width in layers:
        mods += [torch.nn.Linear(prev, width), torch.nn.ReLU()]
        prev = width
    mods.append(torch.nn.Linear(prev, out_len))
    if out_act == "sigmoid":
        mods.append(torch.nn.Sigmoid())
    net = torch.nn.Sequential(*mods)
    for m in net:
        if isinstance(m, torch.nn.Linear):
            torch.nn.init.xavier_uniform_(m.weight, generator=g)
            torch.nn.init.zeros_(m.bias)
    return net.to(device)


class CACLALearner(object):
    """The reference's DEFAULT algorithm (networkParameters.py:1 ALGORITHM = "CACLA"), batched:
    critic V(s) trained on r + discount * V_target(s') (src/model/actorCritic.py:968-983, 1032-1065), actor moved towards
    the taken action where the TD error is positive (:1068-1081), CACLA+Var: ceil(td / sqrt(var)) actor updates per sample
    with the running variance var <- (1 - beta) var + beta td^2 taken sample by sample in batch order (:806-811, 826-840).
    Hyper-parameters: networkParameters.py:141-153 (layers (100, 100), critic lr 7.5e-5, actor lr 5e-4, var start 1,
    beta 1e-3), hard target update every TARGET_NETWORK_STEPS = 1500, sigmoid policy output, Gaussian exploration noise
    1.0 decaying to 0.02 at half training (:55-58), actions clipped to [0, 1] (:922-936).

    One deliberate difference: the reference's epoch loop compacts `inputs` in place while indexing the counts by
    original position (:829-836), so from the second epoch on it trains on shifted rows; this class trains epoch e on
    exactly the samples whose count exceeds e (the intent)."""

    def __init__(self, state_len, action_len=2, layers=(100, 100), critic_lr=0.000075, actor_lr=0.0005, discount=0.90,
                 target_network_steps=1500, var_start=1.0, var_beta=0.001, update_on_negative_td=False, noise=1.0,
                 noise_decay=1.0, max_epochs=None, device="cuda", seed=0):
        import torch
        self.torch = torch
        self.actor = make_mlp(state_len, layers, action_len, "sigmoid", device, seed)
        self.critic = make_mlp(state_len, layers, 1, None, device, seed + 1)
        self.critic_target = copy.deepcopy(self.critic).eval()
        for p in self.critic_target.parameters():
            p.requires_grad_(False)
        self.opt_actor = torch.optim.Adam(self.actor.parameters(), lr=actor_lr)
        self.opt_critic = torch.optim.Adam(self.critic.parameters(), lr=critic_lr)
        self.discount, self.target_network_steps = float(discount), int(target_network_steps)
        self.cacla_var, self.var_beta = float(var_start), float(var_beta)
        self.update_on_negative_td = bool(update_on_negative_td)
        self.std, self.noise_decay = float(noise), float(noise_decay)
        self.max_epochs = max_epochs
        self.gen = torch.Generator(device=device).manual_seed(seed)
        self.step = 0

    # ---- acting (decideMove :938-966 + applyNoise :922-936)
    def decide(self, obs, update_noise=True):
        torch = self.torch
        with torch.no_grad():
            action = self.actor(obs)
            if update_noise:
                self.std *= self.noise_decay
            noisy = action + self.std * torch.randn(action.shape, device=action.device, generator=self.gen)
            return action, noisy.clamp_(0.0, 1.0)

    # ---- learning
    def critic_targets(self, obs_t, reward, obs_tp1, done):
        torch = self.torch
        with torch.no_grad():
            alive = (done == 0).to(torch.float32)
            target = reward.to(torch.float32) + self.discount * self.critic_target(obs_tp1).squeeze(1) * alive
            td = target - self.critic(obs_t).squeeze(1)
        return target, td

    def actor_update_counts(self, td):
        """CACLA+Var: per-sample number of actor updates; advances self.cacla_var through the batch in order."""
        torch = self.torch
        t = td.double()
        n = t.numel()
        keep = 1.0 - self.var_beta
        pw = keep ** torch.arange(1, n + 1, dtype=torch.float64, device=t.device)   # keep^(i+1)
        var = pw * self.cacla_var + self.var_beta * pw * torch.cumsum(t * t / pw, 0)
        self.cacla_var = float(var[-1])
        counts = torch.ceil(t / torch.sqrt(var))
        return torch.where(t > 0, counts, torch.zeros_like(counts)).long(), var

    def learn(self, obs_t, action, reward, obs_tp1, done, weights=None):
        torch = self.torch
        w = torch.ones_like(reward, dtype=torch.float32) if weights is None else weights.to(torch.float32)
        target, td = self.critic_targets(obs_t, reward, obs_tp1, done)
        loss_c = (((self.critic(obs_t).squeeze(1) - target) ** 2) * w).mean()
        self.opt_critic.zero_grad(set_to_none=True)
        loss_c.backward()
        self.opt_critic.step()
        counts, _ = self.actor_update_counts(td)
        pos = (td > 0) & (w != 0)
        a_target = action.to(torch.float32)
        if self.update_on_negative_td:  # target = mu - (a - mu)
            with torch.no_grad():
                mu = self.actor(obs_t)
            a_target = torch.where((td > 0).unsqueeze(1), a_target, 2 * mu - a_target)
            pos = (td != 0) & (w != 0)
            counts = torch.where(td < 0, torch.ones_like(counts), counts)
        n_epochs = int(counts[pos].max().item()) if bool(pos.any()) else 0
        if self.max_epochs is not None:
            n_epochs = min(n_epochs, self.max_epochs)
        for e in range(n_epochs):
            m = pos & (counts > e)
            pred = self.actor(obs_t[m])
            loss_a = ((((pred - a_target[m]) ** 2).mean(dim=1)) * w[m]).mean()
            self.opt_actor.zero_grad(set_to_none=True)
            loss_a.backward()
            self.opt_actor.step()
        self.step += 1
        if self.step % self.target_network_steps == 0:
            self.critic_target.load_state_dict(self.critic.state_dict())
        return td, float(loss_c.detach()), n_epochs
