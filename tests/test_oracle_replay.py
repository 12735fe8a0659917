"""Oracle RandomStack (oracle/replay.py) against the real utils.RandomStack (utils.py:14-146) as
recorded in tests/golden/replay_stack.npz by oracle.make_golden: same seeds -> same accept /
duplicate / eviction decisions after every push and the same augmented get_data batch."""
import random

import numpy as np

from conftest import golden
from oracle.replay import OracleRandomStack, symmetry_cell


def games_of(z):
    offs = np.concatenate([[0], np.cumsum(z["g_off"][:, 1] - z["g_off"][:, 0])])
    out = []
    for g in range(len(z["games"])):
        rec = []
        for i in range(int(offs[g]), int(offs[g + 1])):
            la = tuple(int(x) for x in z["g_last"][i])
            rec.append((str(z["g_states"][i]), z["g_policy"][i], la if la[0] >= 0 else None,
                        float(z["g_value"][i]), np.float32(z["g_weight"][i])))
        out.append((rec, int(z["g_result"][g])))
    return out


def test_push_bookkeeping_and_batch_match_reference():
    z = golden("replay_stack.npz")
    random.seed(int(z["seeds"][0]))
    np.random.seed(int(z["seeds"][1]))
    st = OracleRandomStack(11, length=int(z["length"]))
    for g, (rec, res) in enumerate(games_of(z)):
        assert st.push(list(rec), res) == bool(z["accepted"][g]), g
        assert len(st.data) == z["n_data"][g] and st.black_win == z["black"][g] and st.white_win == z["white"][g], g
        assert len(st.data_len) == z["n_games"][g] and (st.data_len[0] if st.data_len else 0) == z["first_len"][g], g
    assert [r[0] for r in st.data] == [str(s) for s in z["final_states"]]
    assert st.data_len == list(z["final_data_len"]) and st.result == list(z["final_result"])
    assert sum(st.data_len) == len(st.data)
    assert (np.diff(z["n_games"]) == 2).any(), "fixture must contain a colour re-balancing duplicate"
    assert (~z["accepted"]).any(), "fixture must contain a rejected short game"
    b, w, v, p = st.get_data(256)
    assert (b == z["batch_boards"]).all() and (w == z["batch_weights"]).all()
    assert (v == z["batch_values"]).all() and (p == z["batch_policies"]).all()


def test_symmetry_cell_is_the_rot90_flip_point_map():
    S = 7
    for k in range(4):
        for flip in (False, True):
            for (i, j) in [(0, 0), (1, 5), (6, 2), (3, 3)]:
                m = np.zeros((S, S), np.int8)
                m[i, j] = 1
                t = np.rot90(m, k=k, axes=(0, 1))
                if flip:
                    t = np.flip(t, axis=0)
                assert tuple(int(x) for x in np.argwhere(t == 1)[0]) == symmetry_cell(i, j, k, flip, S)
