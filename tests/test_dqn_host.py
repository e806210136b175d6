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
