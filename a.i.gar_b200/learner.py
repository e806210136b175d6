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
        """CACLA+Var: per-sample number of actor updates; advances self.cacla_var through the batch in order.
        The recurrence var_i = keep * var_{i-1} + beta * td_i^2 is evaluated as a scan in chunks of 4096 samples that carry the
        variance across chunks: keep^n inside one chunk stays far from underflow for any beta (a single scan over the whole
        batch divides by keep^i, which underflows from ~7e5 samples on)."""
        torch = self.torch
        t = td.double()
        keep = 1.0 - self.var_beta
        outs, var0 = [], self.cacla_var
        for lo in range(0, t.numel(), 4096):
            c = t[lo:lo + 4096]
            pw = keep ** torch.arange(1, c.numel() + 1, dtype=torch.float64, device=t.device)   # keep^(i+1)
            v = pw * var0 + self.var_beta * pw * torch.cumsum(c * c / pw, 0)
            var0 = float(v[-1])
            outs.append(v)
        var = torch.cat(outs) if outs else t
        assert bool(torch.isfinite(var).all()), "CACLA variance is not finite"
        self.cacla_var = var0
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



def _QNet(state_len, action_len, layers, feed_layer, device, seed):
    """ActionValueNetwork (src/model/actorCritic.py:389-535): Q(s, a) MLP, the action concatenated to the input of hidden layer
    DPG_FEED_ACTION_IN_LAYER (1 = together with the state), relu hidden, linear output, glorot_uniform kernels AND biases."""
    import torch

    class QNet(torch.nn.Module):
        def __init__(self):
            super().__init__()
            g = torch.Generator(device="cpu").manual_seed(seed)
            self.feed = int(feed_layer) - 1
            mods, prev = [], state_len
            for i, width in enumerate(layers):
                if i == self.feed:
                    prev += action_len
                mods.append(torch.nn.Linear(prev, width))
                prev = width
            mods.append(torch.nn.Linear(prev, 1))
            self.layers = torch.nn.ModuleList(mods)
            for m in self.layers:
                torch.nn.init.xavier_uniform_(m.weight, generator=g)
                lim = (6.0 / (m.bias.numel() + 1)) ** 0.5
                torch.nn.init.uniform_(m.bias, -lim, lim, generator=g)

        def forward(self, s, a):
            x = s
            for i, m in enumerate(self.layers[:-1]):
                if i == self.feed:
                    x = torch.cat([x, a], dim=1)
                x = torch.relu(m(x))
            return self.layers[-1](x).squeeze(1)

    return QNet().to(device)


class DPGLearner(object):
    """ALGORITHM = "DPG" (src/model/actorCritic.py:719-789, 986-1030), batched:
    critic  Q(s, a) <- r + discount * Q_target(s', mu_target(s')) while alive (train_critic_DPG), MSE with importance weights;
    actor   through the combined actor -> (frozen) critic model towards Q_target(s, mu_target(s)) + DPG_Q_VAL_INCREASE
            (train_actor_DPG: the reference's way of ascending Q without an explicit policy gradient);
    targets soft updates with DPG_TAU after every step (SOFT_TARGET_UPDATES, softlyUpdateTargetModel :191-199,366-373).
    Hyper-parameters: networkParameters.py:171-190.  Priorities returned = TD errors."""

    def __init__(self, state_len, action_len=2, actor_layers=(100, 100), critic_layers=(100, 100), actor_lr=0.0001,
                 critic_lr=0.0005, discount=0.90, tau=0.001, q_val_increase=2.0, feed_action_in_layer=1, noise=1.0,
                 noise_decay=1.0, actor_is=False, device="cuda", seed=0):
        import torch
        self.torch = torch
        self.actor = make_mlp(state_len, actor_layers, action_len, "sigmoid", device, seed)
        self.actor_target = copy.deepcopy(self.actor)
        self.critic = _QNet(state_len, action_len, critic_layers, feed_action_in_layer, device, seed + 1)
        self.critic_target = copy.deepcopy(self.critic)
        for net in (self.actor_target, self.critic_target):
            for p in net.parameters():
                p.requires_grad_(False)
        self.opt_actor = torch.optim.Adam(self.actor.parameters(), lr=actor_lr)
        self.opt_critic = torch.optim.Adam(self.critic.parameters(), lr=critic_lr)
        self.discount, self.tau, self.q_inc, self.actor_is = float(discount), float(tau), float(q_val_increase), bool(actor_is)
        self.std, self.noise_decay = float(noise), float(noise_decay)
        self.gen = torch.Generator(device=device).manual_seed(seed)
        self.step = 0

    def decide(self, obs, update_noise=True):  # decideMove :938-966 + applyNoise :922-936
        torch = self.torch
        with torch.no_grad():
            action = self.actor(obs)
            if update_noise:
                self.std *= self.noise_decay
            return action, (action + self.std * torch.randn(action.shape, device=action.device, generator=self.gen)).clamp_(0.0, 1.0)

    def critic_step(self, obs_t, action, reward, obs_tp1, done, w):
        torch = self.torch
        with torch.no_grad():
            alive = (done == 0).to(torch.float32)
            target = reward.float() + self.discount * self.critic_target(obs_tp1, self.actor_target(obs_tp1)) * alive
            q_old = self.critic(obs_t, action)
            td = target - q_old
        loss = (((self.critic(obs_t, action) - target) ** 2) * w).mean()
        self.opt_critic.zero_grad(set_to_none=True)
        loss.backward()
        self.opt_critic.step()
        return td, q_old, loss.detach()

    def _soft_update(self):
        torch = self.torch
        with torch.no_grad():
            for tgt, src in ((self.actor_target, self.actor), (self.critic_target, self.critic)):
                for pt, p in zip(tgt.parameters(), src.parameters()):
                    pt.mul_(1.0 - self.tau).add_(p, alpha=self.tau)

    def learn(self, obs_t, action, reward, obs_tp1, done, weights=None):
        torch = self.torch
        w = torch.ones_like(reward, dtype=torch.float32) if weights is None else weights.float()
        action = action.float()
        td, _, loss_c = self.critic_step(obs_t, action, reward, obs_tp1, done, w)
        with torch.no_grad():
            a_target = self.critic_target(obs_t, self.actor_target(obs_t)) + self.q_inc
        for p in self.critic.parameters():  # combinedActorCritic: the critic half is not trainable (:581-598)
            p.requires_grad_(False)
        q_pi = self.critic(obs_t, self.actor(obs_t))
        loss_a = (((q_pi - a_target) ** 2) * (w if self.actor_is else 1.0)).mean()
        self.opt_actor.zero_grad(set_to_none=True)
        loss_a.backward()
        self.opt_actor.step()
        for p in self.critic.parameters():
            p.requires_grad_(True)
        self._soft_update()
        self.step += 1
        return td, float(loss_c), float(loss_a.detach())


class SPGLearner(DPGLearner):
    """ALGORITHM = "SPG" (Sampled Policy Gradient, OCACLA_ENABLED; src/model/actorCritic.py:731-734, 860-919), batched:
    the critic is trained exactly as in DPG; for the actor every sample searches the action space OFFLINE — candidates are the
    stored action (evaluated by the critic before its update, `evals`), the current policy's action and OCACLA_EXPL_SAMPLES
    Gaussian samples (std = ocacla_noise) drawn around the best candidate so far (OCACLA_MOVING_GAUSSIAN) — and the actor is
    regressed (MSE, importance weights) towards the best candidate wherever it beats the current policy's action.
    Hard target updates every TARGET_NETWORK_STEPS (the soft ones are DPG-only, :744-747).  networkParameters.py:159-169."""

    def __init__(self, state_len, action_len=2, expl_samples=5, ocacla_noise=1.0, ocacla_noise_decay=1.0, moving_gaussian=True,
                 target_network_steps=1500, actor_lr=0.0005, critic_lr=0.000075, **kw):
        DPGLearner.__init__(self, state_len, action_len, actor_lr=actor_lr, critic_lr=critic_lr, **kw)
        self.expl_samples, self.ocacla_noise, self.ocacla_noise_decay = int(expl_samples), float(ocacla_noise), float(ocacla_noise_decay)
        self.moving_gaussian, self.target_network_steps = bool(moving_gaussian), int(target_network_steps)

    def learn(self, obs_t, action, reward, obs_tp1, done, weights=None):
        torch = self.torch
        w = torch.ones_like(reward, dtype=torch.float32) if weights is None else weights.float()
        action = action.float()
        td, evals, loss_c = self.critic_step(obs_t, action, reward, obs_tp1, done, w)  # evals = Q(s, a) before the update
        with torch.no_grad():
            mu = self.actor(obs_t)
            q_mu = self.critic(obs_t, mu)
            best_a, best_q = action.clone(), evals.clone()
            better = q_mu > best_q
            best_a = torch.where(better.unsqueeze(1), mu, best_a)
            best_q = torch.where(better, q_mu, best_q)
            for _ in range(self.expl_samples):
                centre = best_a if self.moving_gaussian else mu
                cand = (centre + self.ocacla_noise * torch.randn(centre.shape, device=centre.device, generator=self.gen)).clamp_(0.0, 1.0)
                q_c = self.critic(obs_t, cand)
                better = q_c > best_q
                best_a = torch.where(better.unsqueeze(1), cand, best_a)
                best_q = torch.where(better, q_c, best_q)
            use = best_q > q_mu
        n_used = int(use.sum().item())
        loss_a = 0.0
        if n_used:
            pred = self.actor(obs_t[use])
            la = ((((pred - best_a[use]) ** 2).mean(dim=1)) * w[use]).mean()
            self.opt_actor.zero_grad(set_to_none=True)
            la.backward()
            self.opt_actor.step()
            loss_a = float(la.detach())
        self.ocacla_noise *= self.ocacla_noise_decay
        self.step += 1
        if self.step % self.target_network_steps == 0:
            self.actor_target.load_state_dict(self.actor.state_dict())
            self.critic_target.load_state_dict(self.critic.state_dict())
        updated_actions = torch.where(use.unsqueeze(1), best_a, action)  # written back into the replay buffer by the caller
        return td, float(loss_c), loss_a, updated_actions


class GraphedDQNLoop(object):
    """The whole collector + trainer tick of the reference's Q-learning run (src/aigar.py:844-849 collector, :1090-1215
    trainer, src/model/qLearning.py:146-193) as ONE CUDA graph per tick, nothing leaving the GPU:

        decide (MLP forward, e-greedy with the decayed epsilon of networkParameters.py:118-125) -> agar_step_observe
        (FRAME_SKIP_RATE + 1 frames + the next observation) -> agar_replay_add_batch (transition assembly from the env's own
        device buffers) -> prioritized sample -> batched TD targets / MSE step (Adam, capturable) -> update_priorities

    At 4096 envs the un-graphed loop is launch bound (~100 kernels and a host sync per tick: 1.75e7 env-steps/s in round 1);
    replaying the captured tick removes the launches' CPU cost and every synchronisation.  Host-side work stays outside the
    graph and between replays: the hard target-network update every TARGET_NETWORK_STEPS learner steps, episode resets, logs.
    Epsilon and the tick counter live in device tensors so that the schedule advances inside the graph."""

    def __init__(self, env, net, replay, batch_size=2048, discount=0.90, lr=0.001, target_network_steps=1500,
                 eps_start=1.0, eps_end=0.05, eps_decay_ticks=1800, learn_after=8, num_actions=25):
        import torch
        from .dqn import square_action_table
        self.torch, self.env, self.net, self.replay = torch, env, net, replay
        self.batch_size, self.discount = int(batch_size), float(discount)
        self.target = copy.deepcopy(net).eval()
        for p in self.target.parameters():
            p.requires_grad_(False)
        self.opt = torch.optim.Adam(self.net.parameters(), lr=lr, capturable=True)
        self.target_network_steps, self.learn_after = int(target_network_steps), int(learn_after)
        dev = env.device
        self.table = torch.tensor(square_action_table(num_actions, bool(env.cfg.enable_split), bool(env.cfg.enable_eject)),
                                  dtype=torch.float32, device=dev)
        self.num_actions = self.table.shape[0]
        self.tick_t = torch.zeros((), dtype=torch.float32, device=dev)
        self.eps_start, self.eps_end, self.eps_decay = float(eps_start), float(eps_end), float(eps_decay_ticks)
        self.prev_obs = torch.zeros_like(env.obs)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.period = env.cfg.frame_skip + 1
        self.ticks = 0
        self.steps = 0
        self._graph_collect = None
        self._graph_full = None

    # ---- the two tick bodies (pure device work on torch's current stream)
    def _collect(self):
        torch, env = self.torch, self.env
        E, A, L = env.obs.shape
        with torch.no_grad():
            eps = torch.clamp(self.eps_start - self.tick_t / self.eps_decay * (self.eps_start - self.eps_end), min=self.eps_end)
            q = self.net(env.obs.view(E * A, L))
            idx = q.argmax(dim=1)
            explore = torch.rand(E * A, device=q.device) < eps
            rnd = torch.randint(0, self.num_actions, (E * A,), device=q.device)
            idx = torch.where(explore, rnd, idx)
            self.prev_obs.copy_(env.obs)
            env.step_observe(self.table[idx].view(E, A, 4), self.period)
            self.replay.add_batch(self.prev_obs, idx.float().view(E * A, 1), env.get(lay_GET_REWARD), env.obs,
                                  env.get(lay_GET_DONE), env.get(lay_GET_VALID))
            self.tick_t += 1.0

    def _learn(self):
        torch = self.torch
        u = torch.rand(self.batch_size, dtype=torch.float64, device=self.env.device)
        s, a, r, s2, d, w, ix = self.replay.sample(u)
        a = a.squeeze(1).long().view(-1, 1)
        with torch.no_grad():
            q_old = self.net(s)
            alive = (d == 0).to(q_old.dtype)
            updated = r + self.discount * self.target(s2).max(dim=1).values * alive
            td = updated - q_old.gather(1, a).squeeze(1)
            targets = q_old.clone()
            targets.scatter_(1, a, updated.view(-1, 1))
        pred = self.net(s)
        loss = ((((pred - targets) ** 2).mean(dim=1)) * w.float()).mean()
        self.opt.zero_grad(set_to_none=False)
        loss.backward()
        self.opt.step()
        self.loss.copy_(loss.detach())
        self.replay.update_priorities(ix, td.abs().double() + 1e-4)

    def _capture(self, with_learning):
        torch = self.torch
        dev = self.env.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up outside capture: lazy cuBLAS handles, Adam state, replay scratch allocation
            for _ in range(3):
                self._collect()
                if with_learning:
                    self._learn()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._collect()
            if with_learning:
                self._learn()
        return g, 3

    def run(self, n_ticks):
        """n_ticks collector ticks (each followed by one learner step once the buffer holds `learn_after` ticks)."""
        torch = self.torch
        self.net.train()
        done = 0
        while done < n_ticks:
            if self.ticks < self.learn_after:
                if self._graph_collect is None:
                    self._graph_collect, warm = self._capture(False)
                    self.ticks += warm
                    done += warm
                    continue
                self._graph_collect.replay()
                self.ticks += 1
                done += 1
                continue
            if self._graph_full is None:
                self._graph_full, warm = self._capture(True)
                self.ticks += warm
                self.steps += warm
                done += warm
                continue
            self._graph_full.replay()
            self.ticks += 1
            self.steps += 1
            done += 1
            if self.steps % self.target_network_steps == 0:  # hard update, between replays (qLearning.py:187-193)
                with torch.no_grad():
                    for pt, p in zip(self.target.parameters(), self.net.parameters()):
                        pt.copy_(p)
        return done


from . import layout as _lay  # noqa: E402

lay_GET_REWARD, lay_GET_DONE, lay_GET_VALID = _lay.GET_REWARD, _lay.GET_DONE, _lay.GET_VALID
