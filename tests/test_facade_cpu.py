"""Host-side mirrors of the reference surface that need no GPU: the NetworkAPI pipe protocol
(networkAPI.py:43-78 <-> player.py:194-197), state-string codecs (utils.py:156-196), config
names (config.py:2-27) and the arena / gen_data bookkeeping (choose_best_player.py:59-72,
main.py:86-93)."""
import threading
import time

import numpy as np
import pytest

from conftest import golden
from oracle import rules as orules


class _FakeModel:
    """agent_model stand-in: eval(batch) -> deterministic function of each input."""
    graph = None

    def __init__(self):
        self.batches = []

    def eval(self, x):
        x = np.asarray(x)
        self.batches.append(x.shape[0])
        s = x.reshape(x.shape[0], -1).sum(1)
        p = np.tile(np.arange(121, dtype=np.float32), (x.shape[0], 1)) + s[:, None]
        return p, (s / 1000).astype(np.float32)


def test_networkapi_pipe_protocol():
    from alphafive_b200.genData.networkAPI import NetworkAPI
    model = _FakeModel()
    api = NetworkAPI(None, model)
    pipes = [api.get_pipe() for _ in range(3)]
    api.start(True)
    try:
        # the reference worker: send([x]); spin on poll(); recv()[0] -> (policy f32[C], float)
        for rnd in range(4):
            xs = [np.full((3, 11, 11), float(rnd + 10 * i), np.float32) for i in range(3)]
            for p, x in zip(pipes, xs):
                p.send([x])
            for i, p in enumerate(pipes):
                t0 = time.time()
                while not p.poll():
                    assert time.time() - t0 < 10
                out = p.recv()
                assert isinstance(out, list) and len(out) == 1
                pol, val = out[0]
                s = xs[i].sum()
                assert pol.shape == (121,) and pol.dtype == np.float32 and isinstance(val, float)
                assert pol[5] == np.float32(5 + s) and abs(val - s / 1000) < 1e-3
        # a request carrying several positions gets a reply of the same length, in order
        pipes[0].send([np.zeros((3, 11, 11), np.float32), np.ones((3, 11, 11), np.float32)])
        assert pipes[0].poll(10)
        out = pipes[0].recv()
        assert len(out) == 2 and out[0][1] == 0.0 and abs(out[1][1] - 0.363) < 1e-6
        # a dead worker (EOF) is logged and dropped; the others keep being served
        pipes[2].close()
        time.sleep(0.05)
        pipes[1].send([np.zeros((3, 11, 11), np.float32)])
        assert pipes[1].poll(10) and len(pipes[1].recv()) == 1
        assert sum(model.batches) == 4 * 3 + 2 + 1
    finally:
        api.close()
    api.prediction_worker.join(timeout=5)
    assert not api.prediction_worker.is_alive()


@pytest.mark.parametrize("S", [11, 15])
def test_state_codecs_match_reference_vectors(S):
    from alphafive_b200.genData.player import board_to_state, state_to_board
    g = golden(f"rules_{S}.npz")
    boards = g["boards"][:400]
    for b in boards:
        s = board_to_state(b)
        assert s == orules.encode_state(b)
        assert (state_to_board(s, S) == b).all()
    assert board_to_state(np.zeros((S, S), np.int8)) == (chr(ord("a") + S) + "/") * S


def test_config_module_has_the_reference_names():
    from alphafive_b200 import config
    from alphafive_b200.engine import make_config
    for name, val in dict(board_size=11, simulation_per_step=542, upper_simulation_per_step=642, goal=5, c_puct=5.0,
                          dirichlet_alpha=0.3, gamma=0.94, init_temp=1.2, tau_decay_rate=0.94, tau_decay_rate_r=0.9,
                          max_processes=5, buffer_size=12000, batch_size=512).items():
        assert getattr(config, name) == val
    assert config.get_lr(0) == 1e-3 and config.get_lr(7000) == 2e-4 and config.get_lr(10 ** 9) == 2e-6
    c = make_config(config, n_games=7, training=False, random_a=True)
    assert (c.board_size, c.sims, c.upper_sims, c.n_games, c.training, c.random_a) == (11, 542, 642, 7, 0, 1)
    config_like = type("Cfg", (), dict(board_size=15, simulation_per_step=800, upper_simulation_per_step=900, goal=5))
    c = make_config(config_like, n_games=1)
    assert (c.board_size, c.sims, c.upper_sims) == (15, 800, 900) and abs(c.c_puct - 5.0) < 1e-9


def test_arena_and_gen_data_bookkeeping():
    from alphafive_b200.drivers import count_wins, label_result
    # draws are skipped; from game 30 on the match stops when one side has no win or the ratio leaves [0.5, 2]
    assert count_wins([0, 1, -1, 0]) == (2, 1, 4)
    assert count_wins([0] * 100) == (31, 0, 31)                       # i = 30 is the first check
    assert count_wins([0, 1] * 50) == (50, 50, 100)
    w = [0, 1] * 15 + [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]
    w0, w1, used = count_wins(w)
    assert w0 / w1 > 2.0 and used < len(w) and (w0, w1) == (31, 15)
    rec = lambda L, v: [("s", None, None, v if i == L - 1 else -v, 1.0) for i in range(L)]
    assert label_result(rec(9, 1.0)) == 1 and label_result(rec(10, 1.0)) == -1 and label_result(rec(121, 0.0)) == 0
