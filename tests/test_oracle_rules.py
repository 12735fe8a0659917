"""The numpy oracle (oracle/rules.py) against golden vectors produced by the
reference's own utils.py (oracle/make_golden.py) and the shipped replay buffer."""
import numpy as np
import pytest

from conftest import golden
from oracle import rules


@pytest.mark.parametrize("S", [11, 15])
def test_rules_against_reference_vectors(S):
    g = golden(f"rules_{S}.npz")
    boards = g["boards"]
    for t in range(boards.shape[0]):
        b = boards[t]
        assert rules.terminal_code(b) == g["codes"][t]
        s = rules.encode_state(b)
        assert s == str(g["states"][t])
        assert (rules.decode_state(s, S) == b).all()
        assert rules.legal_cells(b).shape[0] == g["nlegal"][t]
        la = tuple(int(x) for x in g["last_action"][t])
        x = rules.input_planes(b, la if la[0] >= 0 else None)
        assert (x.astype(np.int8) == g["inputs"][t]).all()
        a = tuple(int(v) for v in g["actions"][t])
        if a[0] >= 0:
            assert (rules.play(b, a) == g["stepped"][t]).all()


def test_scan_order_matters_only_for_double_fives():
    b = np.zeros((11, 11), np.int8)
    b[0, 0:5] = -1          # opponent five anchored first in scan order
    b[5, 0:5] = 1
    assert rules.terminal(b) == (True, -1.0)
    assert rules.terminal(-b) == (True, 1.0)
    b2 = np.zeros((11, 11), np.int8)
    b2[0:6, 3] = 1          # overline counts
    assert rules.terminal(b2) == (True, 1.0)
    b3 = np.zeros((11, 11), np.int8)
    b3[7:11, 10] = 1        # clipped run of four at the edge
    assert rules.terminal(b3) == (False, 0.0)


def test_initial_state_strings():
    assert rules.initial_state(11) == "l/" * 11
    assert rules.initial_state(15) == "p/" * 15
    assert rules.encode_state(np.zeros((11, 11), np.int8)) == "l/" * 11


def test_ply_weights_match_reference():
    w = golden("weights.npz")["w"]
    for L in range(1, 65):
        np.testing.assert_allclose(rules.ply_weights(L, 0.94), w[L - 1, :L], rtol=0, atol=1e-6)


def test_replay_buffer_invariants():
    """Shipped self-play data: encoding round trip, step transitions, no stored
    terminal state, alternating value labels, weights."""
    g = golden("replay_sample.npz")
    for s, b in zip(g["states"], g["boards"]):
        assert rules.encode_state(b) == str(s)
        assert rules.terminal_code(b) == 0
    for s, pol, b in zip(g["states"], g["policy"], g["boards"]):
        support = pol.reshape(-1) > 0
        assert not support[b.reshape(-1) != 0].any()      # no mass on stones
    off = g["g_off"]
    for k in range(len(off) - 1):
        sl = slice(off[k], off[k + 1])
        boards, la, val, wt = g["g_boards"][sl], g["g_last_action"][sl], g["g_value"][sl], g["g_weight"][sl]
        L = boards.shape[0]
        for t in range(L - 1):
            move = tuple(int(v) for v in la[t + 1])
            assert (rules.play(boards[t], move) == boards[t + 1]).all()
        assert la[0][0] == -1 and not boards[0].any()
        assert (val[1:] == -val[:-1]).all() and val[-1] == 1.0
        assert int(g["g_result"][k]) == (1 if L % 2 == 1 else -1)
        np.testing.assert_allclose(rules.ply_weights(L, 0.94), wt, atol=1e-6)
        # some move from the final stored position ends the game in favour of the mover
        wins = [rules.terminal(rules.play(boards[-1], a)) for a in rules.legal_actions(boards[-1])]
        assert (True, -1.0) in wins
