"""Host logic of the DQN consumer (BASELINE configs[4]): the reference's discrete action table."""
import pytest

from aigar_b200.dqn import square_action_table


def test_square_action_table_matches_reference_formula():
    t = square_action_table(25)
    assert len(t) == 25
    assert t[0] == [0.1, 0.1, 0, 0] and t[4] == [0.9, 0.1, 0, 0]           # x = col/5 + 0.1 (network.py:55-56)
    assert t[5] == [0.1, 1 / 5 + 0.1, 0, 0] and t[24] == [4 / 5 + 0.1, 4 / 5 + 0.1, 0, 0]
    ts = square_action_table(25, True, True)
    assert len(ts) == 75 and ts[1] == [0.1, 0.1, 1, 0] and ts[2] == [0.1, 0.1, 0, 1]
    with pytest.raises(ValueError):
        square_action_table(24)


def test_cacla_var_counts_equal_the_sequential_loop():
    """CACLALearner.actor_update_counts == the per-sample recurrence of src/model/actorCritic.py:806-811 (CPU torch)."""
    import math
    import numpy as np
    import torch
    from aigar_b200.learner import CACLALearner
    lrn = CACLALearner(10, device="cpu", var_start=1.0, var_beta=0.001)
    rng = np.random.default_rng(3)
    for batch in range(3):
        td = torch.tensor(rng.normal(0, 3, 257), dtype=torch.float32)
        var0 = lrn.cacla_var
        counts, var = lrn.actor_update_counts(td)
        v = var0
        for i, t in enumerate(td.double().tolist()):
            v = (1 - 0.001) * v + 0.001 * (t ** 2)
            assert abs(float(var[i]) - v) < 1e-9 * max(1.0, v)
            want = math.ceil(t / math.sqrt(v)) if t > 0 else 0
            assert int(counts[i]) == want
        assert abs(lrn.cacla_var - v) < 1e-9


def test_cacla_learns_a_bandit():
    """Critic tracks r + discount V(s'), the actor moves towards actions with positive TD error (CPU torch)."""
    import torch
    from aigar_b200.learner import CACLALearner
    torch.manual_seed(0)
    lrn = CACLALearner(4, device="cpu", noise=0.3, critic_lr=0.01, actor_lr=0.01, max_epochs=3)
    s = torch.zeros((256, 4))
    s[:, 0] = 1.0
    for it in range(300):
        _, a = lrn.decide(s)
        r = 1.0 - ((a - torch.tensor([0.8, 0.2])) ** 2).sum(dim=1)   # best action (0.8, 0.2)
        lrn.learn(s, a, r, s, torch.ones(256, dtype=torch.uint8))
    mu, _ = lrn.decide(s[:1], update_noise=False)
    assert abs(float(mu[0, 0]) - 0.8) < 0.1 and abs(float(mu[0, 1]) - 0.2) < 0.1


def test_cacla_var_scan_survives_huge_batches():
    """the single-scan form divided by keep^i, which underflows from ~7e5 samples on; the chunked scan does not."""
    import torch
    from aigar_b200.learner import CACLALearner
    lrn = CACLALearner(4, device="cpu", var_beta=0.01)
    td = torch.randn(900_000)
    counts, var = lrn.actor_update_counts(td)
    assert bool(torch.isfinite(var).all()) and 0.5 < lrn.cacla_var < 2.0 and int(counts.max()) >= 1


def test_dpg_critic_targets_equal_the_per_sample_loop():
    """DPGLearner.critic_step against the per-sample restatement of train_critic_DPG (actorCritic.py:986-1030)."""
    import torch
    from aigar_b200.learner import DPGLearner
    torch.manual_seed(1)
    lrn = DPGLearner(6, device="cpu", critic_lr=0.0)  # lr 0: the step must not move the networks
    B = 17
    s, s2, a = torch.rand(B, 6), torch.rand(B, 6), torch.rand(B, 2)
    r, done = torch.randn(B), (torch.rand(B) < 0.3).to(torch.uint8)
    td, q_old, _ = lrn.critic_step(s, a, r, s2, done, torch.ones(B))
    for i in range(B):
        target = float(r[i])
        if int(done[i]) == 0:
            target += 0.9 * float(lrn.critic_target(s2[i:i + 1], lrn.actor_target(s2[i:i + 1]))[0])
        est = float(lrn.critic(s[i:i + 1], a[i:i + 1])[0].detach())
        assert abs(float(td[i]) - (target - est)) < 1e-5 and abs(float(q_old[i]) - est) < 1e-6


@pytest.mark.parametrize("algo", ["DPG", "SPG"])
def test_dpg_and_spg_learn_a_bandit(algo):
    """Both actor rules find the best action of a one-state problem, reward = 1 - |a - (0.8, 0.2)|^2 (CPU torch)."""
    import torch
    from aigar_b200.learner import DPGLearner, SPGLearner
    torch.manual_seed(0)
    if algo == "DPG":
        lrn = DPGLearner(4, device="cpu", noise=0.3, critic_lr=0.01, actor_lr=0.003, tau=0.05, q_val_increase=0.5)
    else:
        lrn = SPGLearner(4, device="cpu", noise=0.3, critic_lr=0.01, actor_lr=0.003, ocacla_noise=0.3, target_network_steps=50)
    s = torch.zeros((256, 4))
    s[:, 0] = 1.0
    best = torch.tensor([0.8, 0.2])
    for it in range(500):
        _, a = lrn.decide(s)
        r = 1.0 - ((a - best) ** 2).sum(dim=1)
        lrn.learn(s, a, r, s, torch.ones(256, dtype=torch.uint8))  # done: the target is the reward itself
    mu, _ = lrn.decide(s[:1], update_noise=False)
    assert float(((mu[0] - best) ** 2).sum()) < 0.03, mu


def test_spg_updated_actions_are_the_best_candidates():
    """Where the offline search finds an action the critic rates above the current policy's, that action is what the actor is
    regressed to and what is handed back for the replay buffer (train_actor_OCACLA :905-910)."""
    import torch
    from aigar_b200.learner import SPGLearner
    torch.manual_seed(2)
    lrn = SPGLearner(5, device="cpu", expl_samples=3, ocacla_noise=0.2, critic_lr=0.0, actor_lr=0.0)
    B = 64
    s, a = torch.rand(B, 5), torch.rand(B, 2)
    r, done = torch.randn(B), torch.ones(B, dtype=torch.uint8)
    td, lc, la, upd = lrn.learn(s, a, r, s, done)
    with torch.no_grad():
        q_mu = lrn.critic(s, lrn.actor(s))
        q_upd = lrn.critic(s, upd)
        q_a = lrn.critic(s, a)
    changed = (upd != a).any(dim=1)
    assert bool((q_upd[changed] > q_mu[changed]).all())          # replaced only by something better than the policy
    assert bool((q_upd >= torch.minimum(q_a, q_upd) - 1e-6).all()) and upd.min() >= 0 and upd.max() <= 1
