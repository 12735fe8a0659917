"""Device Player against the reference's own recorded game (tests/golden/gui_game_6960.npz, see
tests/test_oracle_gui_game.py): the GUI.py loop with alphafive_b200's Player / ResNet in place of the reference's
must play the AI's 29 moves -- through the one-kernel fp32 latency path (the Player's default for an on-device
pv_fn), through the tcgen05 path, and through the Pipe protocol of NetworkAPI."""
import numpy as np
import pytest

from conftest import golden
from oracle import rules as orules
from test_oracle_gui_game import replay_ai_moves

pytestmark = pytest.mark.gpu


class _Cfg:
    board_size, goal = 11, 5
    simulation_per_step, upper_simulation_per_step = 542, 642
    c_puct, dirichlet_alpha, init_temp, tau_decay_rate, tau_decay_rate_r, gamma = 5.0, 0.3, 1.2, 0.94, 0.9, 0.94
    max_processes = 1


@pytest.mark.parametrize("route", ["small", "tc", "pipe"])
def test_device_player_reproduces_the_reference_ai_moves(cuda_lib, route):
    from alphafive_b200 import _lib
    from alphafive_b200.genData.network import ResNet
    from alphafive_b200.genData.player import Player, board_to_state
    g = golden("gui_game_6960.npz")
    moves = [tuple(int(v) for v in m) for m in g["moves"]]
    z = golden("ckpt6960.npz")
    net = ResNet(11, max_batch=8)
    net.set_weights({k.replace("__", "/"): z[k] for k in z.files})
    cfg = _Cfg()
    if route == "pipe":
        pl = Player(cfg, training=False, pipe=net.get_pipes(cfg))
    else:
        pl = Player(cfg, training=False, pv_fn=net.eval)
        pl.net_mode = _lib.NET_SMALL if route == "small" else _lib.NET_TC

    def get_action(board, last):
        policy, action = pl.get_action(board_to_state(board), last_action=last, random_a=False)     # GUI.py:154
        assert policy is None
        pl.pruning_tree(board, board_to_state(orules.play(board, action)))                         # GUI.py:160
        return action

    got = replay_ai_moves(get_action, moves)
    assert got == moves[0::2], (route, [(2 * i, a, b) for i, (a, b) in enumerate(zip(got, moves[0::2])) if a != b])
    pl.close()
    net.close()
