"""The reference-facing surface on the GPU: utils free functions, ResNet / NetworkAPI / Player
objects and the three driver loops, checked against the oracle."""
import types

import numpy as np
import pytest

from conftest import golden
from oracle import mcts as omcts, net as onet, rules as orules

pytestmark = pytest.mark.gpu


def _cfg(**kw):
    from alphafive_b200 import config
    ns = types.SimpleNamespace(**{k: getattr(config, k) for k in dir(config) if not k.startswith("_")})
    for k, v in kw.items():
        setattr(ns, k, v)
    return ns


def test_utils_free_functions(cuda_lib):
    from alphafive_b200 import utils
    g = golden("rules_11.npz")
    rng = np.random.default_rng(0)
    for b in g["boards"][rng.choice(len(g["boards"]), 60, replace=False)]:
        over, val = utils.is_game_over(b, 5)
        assert (over, val) == orules.terminal(b) and isinstance(over, bool) and isinstance(val, float)
        legal = utils.get_legal_actions(b)
        assert legal == orules.legal_actions(b)
        la = legal[len(legal) // 2] if legal else None
        x = utils.board_to_inputs(b, last_action=la)
        assert x.dtype == np.float32 and (x == orules.input_planes(b, la)).all()
        if legal:
            b2 = b.copy()
            nxt = utils.step(b2, la)
            assert b2[la] == 1 and (nxt == orules.play(b, la)).all()          # in place + negated copy
        assert utils.state_to_board(utils.board_to_state(b), 11).tolist() == b.tolist()
    assert (utils.BLACK_WIN, utils.WHITE_WIN, utils.DRAW) == (1, -1, 0)
    np.testing.assert_allclose(utils.construct_weights(30, 0.94), orules.ply_weights(30, 0.94), atol=1e-6)


def test_resnet_surface_and_pipe_mode(cuda_lib):
    """ResNet.eval is the pv_fn seam; get_pipes serves Player(pipe=...) through NetworkAPI; the
    two leaf-evaluation routes give the same deterministic search."""
    from alphafive_b200.genData.network import ResNet
    from alphafive_b200.genData.player import Player
    z = golden("ckpt6960.npz")
    w = {k.replace("__", "/"): z[k] for k in z.files}
    net = ResNet(11, max_batch=64)
    net.set_weights(w)
    g = golden("replay_sample.npz")
    x = np.stack([orules.input_planes(b, tuple(la) if la[0] >= 0 else None) for b, la in zip(g["boards"][:32], g["last_action"][:32])])
    p, v = net.eval(x)
    op, ov = onet.OracleNet(11, w).eval(x)
    assert np.abs(p - op).max() < 1e-4 and np.abs(v - ov).max() < 1e-4
    assert np.abs(net.get_prob(x) - p).max() == 0 and np.abs(net.get_value(x) - v).max() == 0
    with pytest.raises(FileNotFoundError):
        net.restore("/nonexistent/ckpt")
    cfg = _cfg(simulation_per_step=40, upper_simulation_per_step=60)
    state = "l/" * 5 + "e31f/" + "l/" * 5
    a_fn = Player(cfg, training=False, pv_fn=net.eval)
    pol, act = a_fn.get_action(state, last_action=(5, 5))
    assert pol is None and act in orules.legal_actions(orules.decode_state(state, 11))
    pipe = net.get_pipes(cfg)
    a_pipe = Player(cfg, training=False, pipe=pipe)
    _, act2 = a_pipe.get_action(state, last_action=(5, 5))
    assert act2 == act
    # the same search in the oracle with the oracle net (1e-4-close leaf values): same move
    opl = omcts.OraclePlayer(omcts.SearchConfig(simulation_per_step=40, upper_simulation_per_step=60), training=False,
                             pv_fn=onet.OracleNet(11, w).eval)
    _, oact = opl.get_action(orules.decode_state(state, 11), (5, 5))
    assert tuple(oact) == tuple(act)
    assert len(a_fn.tree) > 30 and state in a_fn.tree
    a_fn.close(); a_pipe.close(); net.close()


def test_player_training_mode_and_run(cuda_lib):
    from alphafive_b200.genData.network import ResNet
    from alphafive_b200.genData.player import Player
    from alphafive_b200.drivers import label_result
    cfg = _cfg(simulation_per_step=30, upper_simulation_per_step=40)
    net = ResNet(11, max_batch=8)
    pl = Player(cfg, training=True, pv_fn=net.eval, seed=3)
    state = pl.get_init_state()
    assert state == "l/" * 11
    pol, act = pl.get_action(state)
    assert pol.shape == (11, 11) and pol.dtype == np.float32 and abs(pol.sum() - 1) < 1e-4 and pol[act] > 0
    assert abs(pl.tau - cfg.init_temp * cfg.tau_decay_rate) < 1e-12
    pl.reset()
    assert pl.tau == cfg.init_temp and pl.tree == {}
    rec = pl.run()
    L = len(rec)
    assert 9 <= L <= 121 and rec[0][0] == "l/" * 11 and rec[0][2] is None
    for t in range(L - 1):
        b = orules.decode_state(rec[t][0], 11)
        assert orules.encode_state(orules.play(b, rec[t + 1][2])) == rec[t + 1][0]
    vals = [r[3] for r in rec]
    assert vals[-1] in (1.0, 0.0) and all(vals[i + 1] == -vals[i] for i in range(L - 1))
    np.testing.assert_allclose([r[4] for r in rec], orules.ply_weights(L, 0.94), atol=1e-6)
    assert label_result(rec) in (1, -1, 0)
    pl.close(); net.close()


def test_driver_loops(cuda_lib):
    """self_play.py:94-101 viewer game, and the lock-step arena of choose_best_player.py:42-72."""
    from alphafive_b200.drivers import Arena, count_wins, self_play_game
    from alphafive_b200.net import DeviceNet, glorot_init
    from alphafive_b200.genData.network import ResNet
    cfg = _cfg(simulation_per_step=24, upper_simulation_per_step=34)
    net = ResNet(11, max_batch=8)
    moves, state, value = self_play_game(cfg, net.eval, max_plies=12)
    assert len(moves) == 12 and len(set(moves)) == 12
    b = orules.decode_state(state, 11)
    assert np.count_nonzero(b) == 12 and abs(int(b.sum())) == 0
    net.close()
    N = 16
    n0, n1 = DeviceNet(11, N, glorot_init(11, 0)), DeviceNet(11, N, glorot_init(11, 1))
    arena = Arena(cfg, n0, n1, N, seed=5)
    out = arena.play()
    assert set(np.unique(out["winners"])) <= {-1, 0, 1} and (out["plies"] >= 9).all() and (out["plies"] <= 121).all()
    finals = [orules.decode_state(s, 11) for s in out["final_states"]]
    for i, fb in enumerate(finals):
        over, val = orules.terminal(fb)
        assert over and np.count_nonzero(fb) == out["plies"][i]
        if val == 0.0:
            assert out["winners"][i] == -1
        else:   # the side to move has lost; the mover of the last ply is (first + plies - 1) % 2
            assert val == -1.0 and out["winners"][i] == (i + out["plies"][i] - 1) % 2
    w0, w1, used = count_wins(out["winners"].tolist())
    assert w0 + w1 <= N and used == N
    assert out["moves"] == int(out["plies"].sum())
    arena.close()


def test_player_tree_view_pruning_and_reset(cuda_lib):
    """Player.tree (player.py:29) as a mapping of State objects, equal to the oracle's table; the device GC
    leaves no key the reference's pruning_tree (player.py:149-164) would delete; reset(search_tree=...)."""
    from alphafive_b200.genData.player import Player, state_to_board, board_to_state
    S = 11
    cfg = _cfg(simulation_per_step=90, upper_simulation_per_step=120)
    pv = omcts.table_pv_fn(S, 5)
    pl = Player(cfg, training=False, pv_fn=pv)
    ocfg = omcts.SearchConfig(simulation_per_step=90, upper_simulation_per_step=120)
    opl = omcts.OraclePlayer(ocfg, training=False, pv_fn=pv)
    assert len(pl.tree) == 0
    state, last = pl.get_init_state(), None
    for ply in range(4):
        board = state_to_board(state, S)
        _, action = pl.get_action(state, last_action=last)
        _, oaction = opl.get_action(board, last)
        assert action == oaction
        tree = pl.tree
        # every key the oracle's (never pruned) table holds *and* that contains the root's stones is there,
        # with the same sum_n; nothing else is
        root_own, root_opp = board == 1, board == -1
        want = {}
        for key, node in opl.table.items():
            b = np.frombuffer(key, np.int8).reshape(S, S)
            q = int((b != 0).sum()) - int((board != 0).sum())
            own, opp = (b == 1, b == -1) if q % 2 == 0 else (b == -1, b == 1)
            if q >= 0 and (own >= root_own).all() and (opp >= root_opp).all():
                want[board_to_state(b)] = node.sum_n
        assert {k: v.sum_n for k, v in tree.items()} == want
        # reference pruning predicate relative to the current root: nothing left to delete
        for key in tree:
            b = state_to_board(key, S)
            deletable = key != state and (root_own >= (b == 1)).all() and (root_opp >= (b == -1)).all()
            assert not deletable, key
        assert pl.pruning_tree(board, state) is None and len(pl.tree) == len(want)
        # a node as the reference exposes it
        node = tree[state]
        n, w, p, sum_n = opl.root_stats(board)
        assert node.sum_n == sum_n and set(node.a) == set(orules.legal_actions(board))
        for (i, j), e in node.a.items():
            c = i * S + j
            assert e.n == n[c] and e.w == w[c] and e.p == p[c] and (e.q == w[c] / np.float32(n[c]) if n[c] else e.q == 0)
        with pytest.raises(KeyError):
            tree["x" + state]
        nxt = orules.play(board, action)
        state, last = board_to_state(nxt), action
    with pytest.raises(NotImplementedError):
        pl.reset(search_tree={})
    pl.reset(None)
    assert len(pl.tree) == 0 and pl.root_state is None
    pl.close()


def test_player_engine_follows_budget_growth(cuda_lib):
    """config.simulation_per_step is read lazily and may grow (choose_best_player.py:25): the table is
    re-sized instead of silently dropping expansions."""
    from alphafive_b200.genData.player import Player
    cfg = _cfg(simulation_per_step=20, upper_simulation_per_step=30)
    pv = omcts.table_pv_fn(11, 2)
    pl = Player(cfg, training=False, pv_fn=pv)
    s0 = pl.get_init_state()
    pl.get_action(s0)
    cap0 = pl._engine.cfg.node_capacity
    cfg.simulation_per_step, cfg.upper_simulation_per_step = 3000, 3100
    _, action = pl.get_action(s0)
    assert pl._engine.cfg.node_capacity > cap0
    opl = omcts.OraclePlayer(omcts.SearchConfig(simulation_per_step=3000, upper_simulation_per_step=3100),
                             training=False, pv_fn=pv)
    assert opl.get_action(np.zeros((11, 11), np.int8), None)[1] == action
    assert pl.tree[s0].sum_n == 2999                     # the first simulation expands the root, the other 2999 select from it
    pl.close()
