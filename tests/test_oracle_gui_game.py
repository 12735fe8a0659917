"""The one first-party known answer of network + search together: the human-vs-AI game the reference ships as
tmp/five_6960.gif (written by GUI.py:184-186; decoded by oracle/make_golden.py:make_gui_game into
tests/golden/gui_game_6960.npz).  The AI (ckpt-6960 through the TensorFlow net, Player(training=False), 542 / 642
simulations from config.py:4-5, one search table kept for the whole game) moved first; its 29 moves are a
deterministic function of the checkpoint, the search and the human's replies.  The oracle -- the torch restatement of
network.py:52-88 driving the restatement of player.py -- must reproduce every one of them, including the close calls
(ply 32: 198 against 193 visits)."""
import numpy as np

from conftest import golden
from oracle import mcts as omcts, net as onet, rules as orules


def replay_ai_moves(get_action, moves):
    """GUI.py:108-167: AI to move on even plies with last_action = the human's previous move (None at the start);
    the human's moves are replayed from the recording.  Returns the AI's choices."""
    board, last, out = np.zeros((11, 11), np.int8), None, []
    for k, mv in enumerate(moves):
        if k % 2 == 0:
            out.append(tuple(int(v) for v in get_action(board, last)))
        board = orules.play(board, mv)
        last = mv
    assert orules.terminal(board) == (True, -1.0)
    return out


def test_oracle_reproduces_the_reference_ai_moves():
    g = golden("gui_game_6960.npz")
    moves = [tuple(int(v) for v in m) for m in g["moves"]]
    z = golden("ckpt6960.npz")
    w = {k.replace("__", "/"): z[k] for k in z.files}
    cfg = omcts.SearchConfig(simulation_per_step=int(g["sims"]), upper_simulation_per_step=int(g["upper"]))
    pl = omcts.OraclePlayer(cfg, training=False, pv_fn=onet.OracleNet(11, w).eval)
    got = replay_ai_moves(lambda b, la: pl.get_action(b, la)[1], moves)
    assert got == moves[0::2], [(2 * i, a, b) for i, (a, b) in enumerate(zip(got, moves[0::2])) if a != b]
    assert len(got) == 29
