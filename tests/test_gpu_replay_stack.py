"""Device-resident RandomStack (alphafive_b200.replay_stack, a5_replay_sample) against the real
utils.RandomStack (tests/golden/replay_stack.npz) and the oracle restatement: bit-exact batches."""
import random

import numpy as np
import pytest
import torch

from conftest import golden
from test_oracle_replay import games_of

pytestmark = pytest.mark.gpu


def test_same_seeds_same_decisions_and_batch_as_reference(cuda_lib):
    from alphafive_b200.utils import RandomStack
    z = golden("replay_stack.npz")
    random.seed(int(z["seeds"][0]))
    np.random.seed(int(z["seeds"][1]))
    st = RandomStack(11, length=int(z["length"]))
    for g, (rec, res) in enumerate(games_of(z)):
        assert st.push(list(rec), res) == bool(z["accepted"][g]), g
        assert len(st) == z["n_data"][g] and st.black_win == z["black"][g] and st.white_win == z["white"][g], g
        assert len(st.data_len) == z["n_games"][g] and (st.data_len[0] if st.data_len else 0) == z["first_len"][g], g
    assert st.is_full() and not st.isEmpty()
    held = st.data
    assert [r[0] for r in held] == [str(s) for s in z["final_states"]]
    assert st.data_len == list(z["final_data_len"]) and st.result == list(z["final_result"])
    b, w, v, p = st.get_data(256)
    assert b.dtype == np.float32 and b.shape == (256, 3, 11, 11) and p.shape == (256, 121)
    assert (b == z["batch_boards"]).all() and (w == z["batch_weights"]).all()
    assert (v == z["batch_values"]).all() and (p == z["batch_policies"]).all()


@pytest.mark.parametrize("S", [11, 15])
def test_all_symmetries_against_oracle_on_engine_records(cuda_lib, S):
    """Records harvested from lock-step self-play, every (rotation, flip) pair, ring wrapped around."""
    from alphafive_b200.net import DeviceNet, glorot_init
    from alphafive_b200.replay import parse_records, records_to_games
    from alphafive_b200.selfplay import SelfPlay
    from alphafive_b200.utils import RandomStack
    from oracle.replay import OracleRandomStack
    N = 64
    net = DeviceNet(S, N, glorot_init(S, 0))
    sp = SelfPlay(None, n_games=N, net=net, training=True, seed=3, board_size=S, simulation_per_step=12,
                  upper_simulation_per_step=16)
    sp.start()
    buf = None
    for _ in range(40):
        sp.run_passes(60)
        b, games = sp.harvest()
        if b.shape[0]:
            buf = b.clone() if buf is None else torch.cat([buf, b])
        if buf is not None and buf.shape[0] > 700:
            break
    assert buf is not None and buf.shape[0] > 200
    random.seed(5)
    np.random.seed(6)
    dev = RandomStack(S, length=300)
    flags = dev.push_records(buf)
    random.seed(5)
    np.random.seed(6)
    ora = OracleRandomStack(S, length=300)
    # the engine emits whole games contiguously in ply order: cut them on (game_id, game_serial) runs
    recs = parse_records(buf, S)
    i, oflags = 0, []
    while i < len(recs):
        n = recs[i]["game_len"]
        game, res = records_to_games(recs[i:i + n], S)[0]
        oflags.append(ora.push(list(game), res))
        i += n
    assert flags == oflags and len(dev) == len(ora.data) and dev.data_len == ora.data_len and dev.result == ora.result
    assert dev.head != 0, "the ring should have wrapped / advanced in this test"
    rng = np.random.default_rng(0)
    num = 8 * 40
    idx = rng.integers(0, len(dev), num)
    rot = (np.arange(num) % 4).astype(np.uint8)
    flip = ((np.arange(num) // 4) % 2).astype(np.uint8)
    b, w, v, p = dev.gather((dev.head + idx) % dev.capacity, rot, flip)
    ob, ow, ov, op = ora.gather(idx, rot, flip.astype(bool))
    assert (b.cpu().numpy() == ob).all() and (p.cpu().numpy() == op).all()
    assert (w.cpu().numpy() == ow).all() and (v.cpu().numpy() == ov).all()
    assert b.cpu().numpy()[:, 2].sum(axis=(1, 2)).max() <= 1.0


def test_pickle_interchange_with_reference_format(cuda_lib, tmp_path):
    from alphafive_b200.utils import RandomStack
    from oracle.replay import OracleRandomStack
    z = golden("replay_stack.npz")
    random.seed(1)
    st = RandomStack(11, length=400)
    for rec, res in games_of(z)[:30]:
        st.push(list(rec), res)
    st.save("7", directory=str(tmp_path))
    import pickle
    data = pickle.load(open(tmp_path / "data7.pkl", "rb"))
    assert isinstance(data, list) and len(data) == len(st) and isinstance(data[0][0], str)
    assert data[0][1].shape == (11, 11) and data[0][1].dtype == np.float32 and isinstance(data[0][4], np.float32)
    st2 = RandomStack(11, length=400)
    st2.load("7", directory=str(tmp_path))
    assert len(st2) == len(st) and st2.data_len == st.data_len and st2.result == st.result
    assert st2.black_win == st.black_win and st2.white_win == st.white_win
    ora = OracleRandomStack(11, length=400)
    ora.data, ora.data_len, ora.result = data, list(st.data_len), list(st.result)
    idx, rot, flip = np.arange(len(st)), np.arange(len(st)) % 4, (np.arange(len(st)) // 4) % 2 == 1
    b, w, v, p = st2.gather((st2.head + idx) % st2.capacity, rot.astype(np.uint8), flip.astype(np.uint8))
    ob, ow, ov, op = ora.gather(idx, rot, flip)
    assert (b.cpu().numpy() == ob).all() and (p.cpu().numpy() == op).all() and (v.cpu().numpy() == ov).all()
