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
