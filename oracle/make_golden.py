"""Oracle (test infrastructure): write tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container only (``python -m oracle.make_golden``): it imports
``utils``, ``config`` and ``genData.player.Player`` from ``/root/reference`` and
records their outputs on seeded inputs.  ``/root/reference`` does not exist on the
GPU box, so the tests read only the committed ``.npz`` files.

Fixtures
--------
rules_{S}.npz        boards + what utils.is_game_over / board_to_state / step /
                     get_legal_actions / board_to_inputs return for them
mcts_kat_{S}.npz     deterministic Player(training=False) root statistics after k
                     simulations under oracle.mcts.table_pv_fn (tie-free, checked)
mcts_game_{S}.npz    a multi-move deterministic game with tree reuse (budget rule
                     player.py:140-143, retained sub-tree statistics)
mcts_train_11.npz    seeded training-mode summary statistics (distributional pins)
replay_sample.npz    1,024 records + 3 whole games of data_buffer/data6960.pkl
ckpt6960.npz         the 42 tensors of ckpt/alphaFive-6960 (via alphafive_b200.ckpt)
ckpt6960_files.npz   the shipped .index file verbatim + sha256 of the .data file (pins the writer)
weights.npz          construct_weights(L, 0.94) for L = 1..64
policy_vectors.npz   calc_policy outputs (policy, tau) of the real Player over consecutive moves
mcts_mix_11.npz      root / depth-1 visit counts of seeded training-mode searches (mix weights)
selfplay_games_11.npz lengths / results of seeded Player.run() games (training mode)
replay_stack.npz     utils.RandomStack driven with seeded generators: accept flags and
                     bookkeeping after every push, one get_data batch
buffer_games_6960.npz lengths / results of the 470 games of the shipped replay buffer
gui_game_6960.npz    the 58 moves of the human-vs-AI game the reference ships as tmp/five_6960.gif (written by
                     GUI.py:184-186 with the TensorFlow net of ckpt-6960): the 29 AI moves are the one first-party
                     known answer of network + search together
"""
from __future__ import annotations

import os
import pickle
import random
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _ref():
    sys.path.insert(0, REF)
    import config  # noqa
    import utils  # noqa
    from genData.player import Player  # noqa
    return utils, config, Player


def _code(over, v):
    return 0 if not over else 1 if v > 0 else 2 if v < 0 else 3


def adversarial_boards(S, rng):
    """Fives of both colours, overlines, edge-clipped runs, full boards."""
    out = []
    for _ in range(300):
        b = np.zeros((S, S), np.int8)
        for _ in range(rng.integers(1, 4)):
            colour = rng.choice([-1, 1])
            length = int(rng.integers(4, 8))
            d = [(1, 0), (0, 1), (1, 1), (-1, 1)][rng.integers(0, 4)]
            i, j = int(rng.integers(0, S)), int(rng.integers(0, S))
            for k in range(length):
                ii, jj = i + d[0] * k, j + d[1] * k
                if 0 <= ii < S and 0 <= jj < S:
                    b[ii, jj] = colour
        noise = rng.random((S, S))
        b[(b == 0) & (noise < 0.15)] = 1
        b[(b == 0) & (noise > 0.85)] = -1
        out.append(b)
    for _ in range(60):                                   # full / nearly full boards
        b = rng.choice(np.array([-1, 1], np.int8), size=(S, S))
        # checker-ish pattern breaks most fives so that draws actually occur
        if rng.random() < 0.7:
            base = np.fromfunction(lambda i, j: ((i // 2 + j) % 2) * 2 - 1, (S, S)).astype(np.int8)
            flip = rng.random((S, S)) < 0.03
            b = np.where(flip, -base, base).astype(np.int8)
        if rng.random() < 0.5:
            b[rng.integers(0, S), rng.integers(0, S)] = 0
        out.append(b)
    return out


def make_rules(S, utils, n_random, seed):
    from oracle import rules
    rng = np.random.default_rng(seed)
    boards = [rules.random_board(rng, S) for _ in range(n_random)]
    boards += adversarial_boards(S, rng)
    boards.append(np.zeros((S, S), np.int8))
    boards = np.stack(boards).astype(np.int8)
    N = boards.shape[0]
    codes = np.zeros(N, np.int8)
    states = []
    acts = np.zeros((N, 2), np.int16)
    last = np.full((N, 2), -1, np.int16)
    stepped = np.zeros_like(boards)
    inputs = np.zeros((N, 3, S, S), np.int8)
    nlegal = np.zeros(N, np.int16)
    for t in range(N):
        b = boards[t]
        codes[t] = _code(*utils.is_game_over(b.copy(), 5))
        s = utils.board_to_state(b)
        assert (utils.state_to_board(s, S) == b).all()
        states.append(s)
        legal = utils.get_legal_actions(b)
        nlegal[t] = len(legal)
        la = None
        if rng.random() < 0.8:
            la = (int(rng.integers(0, S)), int(rng.integers(0, S)))
            last[t] = la
        inputs[t] = utils.board_to_inputs(b, last_action=la).astype(np.int8)
        if legal:
            a = legal[int(rng.integers(0, len(legal)))]
            acts[t] = a
            stepped[t] = utils.step(b.copy(), a)
        else:
            acts[t] = (-1, -1)
            stepped[t] = b
    np.savez_compressed(os.path.join(OUT, f"rules_{S}.npz"), boards=boards, codes=codes,
                        states=np.array(states), actions=acts, stepped=stepped,
                        last_action=last, inputs=inputs, nlegal=nlegal)
    print(f"rules_{S}: {N} boards; codes", np.bincount(codes, minlength=4))


class _TieWatch:
    """Wraps np.random.choice / random.choice to prove a run never had to break a tie."""

    def __init__(self):
        self.worst = 1          # ties inside the search (np.random.choice, player.py:278)
        self.worst_final = 1    # ties in the final most-visited pick (random.choice, player.py:102)

    def __enter__(self):
        self._np, self._py = np.random.choice, random.choice

        def np_choice(a, *args, **kw):
            if kw.get("p") is None and not args:
                self.worst = max(self.worst, len(a) if hasattr(a, "__len__") else 1)
            return self._np(a, *args, **kw)

        def py_choice(seq):
            self.worst_final = max(self.worst_final, len(seq))
            return self._py(seq)

        np.random.choice, random.choice = np_choice, py_choice
        return self

    def __exit__(self, *exc):
        np.random.choice, random.choice = self._np, self._py


def _root_arrays(player, state, S):
    node = player.tree[state]
    n = np.zeros(S * S, np.int32)
    w = np.zeros(S * S, np.float32)
    p = np.zeros(S * S, np.float32)
    for (i, j), e in node.a.items():
        n[i * S + j], w[i * S + j], p[i * S + j] = e.n, e.w, e.p
    return n, w, p, node.sum_n


def make_mcts_kat(S, utils, config, Player, roots, ks, salt):
    from oracle.mcts import table_pv_fn
    config.board_size = S
    pv = table_pv_fn(S, salt)
    rec = dict(root_boards=[], root_last=[], k=[], n=[], w=[], p=[], sum_n=[], action=[], nkeys=[],
               key_boards=[], key_sum_n=[], key_off=[0])
    for board, la in roots:
        for k in ks:
            config.simulation_per_step = k
            config.upper_simulation_per_step = k + 100
            pl = Player(config, training=False, pv_fn=pv)
            state = utils.board_to_state(board)
            with _TieWatch() as tw:
                _, action = pl.get_action(state, last_action=la)
            assert tw.worst == 1, f"tie encountered (S={S}, k={k})"
            if tw.worst_final > 1:          # e.g. k = 1: no edge visited yet, the pick is random
                action = (-1, -1)
            n, w, p, sum_n = _root_arrays(pl, state, S)
            rec["root_boards"].append(board); rec["root_last"].append(la if la else (-1, -1))
            rec["k"].append(k); rec["n"].append(n); rec["w"].append(w); rec["p"].append(p)
            rec["sum_n"].append(sum_n); rec["action"].append(action); rec["nkeys"].append(len(pl.tree))
            keys = sorted(pl.tree.keys())
            rec["key_boards"].extend(utils.state_to_board(s, S) for s in keys)
            rec["key_sum_n"].extend(pl.tree[s].sum_n for s in keys)
            rec["key_off"].append(len(rec["key_boards"]))
    np.savez_compressed(os.path.join(OUT, f"mcts_kat_{S}.npz"), salt=salt,
                        **{k: np.asarray(v) for k, v in rec.items()})
    print(f"mcts_kat_{S}: {len(rec['k'])} cases, {len(rec['key_boards'])} table keys")


def make_mcts_game(S, utils, config, Player, sims, upper, plies, salt):
    """Tries successive salts until the whole game is tie-free (visit-count ties in the
    final pick are common at low simulation counts)."""
    from oracle.mcts import table_pv_fn
    config.board_size = S
    config.simulation_per_step, config.upper_simulation_per_step = sims, upper
    for salt in range(salt, salt + 200):
        pl = Player(config, training=False, pv_fn=table_pv_fn(S, salt))
        state, action = pl.get_init_state(), None
        rec = dict(boards=[], last=[], n=[], w=[], sum_n=[], action=[], budget=[], nkeys=[])
        with _TieWatch() as tw:
            for _ in range(plies):
                seen = state in pl.tree
                budget = sims if not seen else min(sims, upper - pl.tree[state].sum_n)
                la = action
                _, action = pl.get_action(state, last_action=la)
                n, w, _, sum_n = _root_arrays(pl, state, S)
                board = utils.state_to_board(state, S)
                rec["boards"].append(board.copy()); rec["last"].append(la if la else (-1, -1))
                rec["n"].append(n); rec["w"].append(w); rec["sum_n"].append(sum_n)
                rec["action"].append(action); rec["budget"].append(budget); rec["nkeys"].append(len(pl.tree))
                board = utils.step(board, action)
                state = utils.board_to_state(board)
                if utils.is_game_over(board, 5)[0] or tw.worst > 1 or tw.worst_final > 1:
                    break
        if tw.worst == 1 and tw.worst_final == 1:
            break
    else:
        raise AssertionError("no tie-free salt found for the game KAT")
    np.savez_compressed(os.path.join(OUT, f"mcts_game_{S}.npz"), salt=salt, sims=sims, upper=upper,
                        **{k: np.asarray(v) for k, v in rec.items()})
    print(f"mcts_game_{S}: salt {salt}, {len(rec['action'])} plies, budgets {rec['budget']}")


def make_mcts_train(utils, config, Player, salt):
    """Seeded summary statistics of training-mode searches (Dirichlet noise at every
    node visit, forced-visit ladder at the root): distributional pins."""
    from oracle.mcts import table_pv_fn
    S = 11
    config.board_size = S
    config.simulation_per_step, config.upper_simulation_per_step = 300, 400
    np.random.seed(1234); random.seed(1234)
    runs = 64
    ns = np.zeros((runs, S * S), np.int32)
    depth = np.zeros(runs)
    for r in range(runs):
        pl = Player(config, training=True, pv_fn=table_pv_fn(S, salt))
        state = pl.get_init_state()
        pl.root_state = state                     # what get_action does first (player.py:138)
        for _ in range(300):
            pl.MCTS_search(state, [state], None)
        ns[r] = _root_arrays(pl, state, S)[0]
        depth[r] = sum(st.sum_n for st in pl.tree.values()) / 300.0
    np.savez_compressed(os.path.join(OUT, "mcts_train_11.npz"), salt=salt, sims=300, n=ns, depth=depth)
    print("mcts_train_11: min visits", ns.min(), "mean max", ns.max(1).mean(), "depth", depth.mean())



def make_policy_vectors(utils, config, Player):
    """calc_policy vectors (player.py:84-126) from the real Player.

    det_*   : Player(training=False).get_action(random_a=True) move after move under the tie-free
              table pv_fn -- the search is deterministic, so a device engine fed the same roots
              holds the same visit counts and must return the same policy and tau.  Two sequences:
              init_temp = 1.2 (soft branch) and init_temp = 0.02 (crosses tau <= 0.01 at call 7).
    train_* : Player(training=True) seeded: (n, tau) -> policy vectors of the soft branch with
              noisy visit counts (checked against oracle.mcts.soft_policy on the CPU; the device is
              compared with that formula on its own counts)."""
    from oracle.mcts import table_pv_fn
    S = 11
    config.board_size = S
    out = {}
    keep = (config.simulation_per_step, config.upper_simulation_per_step, config.init_temp)
    config.simulation_per_step, config.upper_simulation_per_step = 120, 135
    for tag, temp, plies in (("det_hi", 1.2, 14), ("det_lo", 0.02, 12)):
        config.init_temp = temp
        for salt in range(30, 230):
            np.random.seed(77 + salt); random.seed(77 + salt)
            pl = Player(config, training=False, pv_fn=table_pv_fn(S, salt))
            state, action = pl.get_init_state(), None
            rec = dict(boards=[], last=[], n=[], policy=[], tau=[], action=[], budget=[])
            ok = True
            with _TieWatch() as tw:
                for _ in range(plies):
                    seen = state in pl.tree
                    budget = 120 if not seen else min(120, 135 - pl.tree[state].sum_n)
                    la = action
                    policy, action = pl.get_action(state, last_action=la, random_a=True)
                    n = _root_arrays(pl, state, S)[0]
                    board = utils.state_to_board(state, S)
                    rec["boards"].append(board.copy()); rec["last"].append(la if la else (-1, -1))
                    rec["n"].append(n); rec["policy"].append(policy.reshape(-1).copy()); rec["tau"].append(pl.tau)
                    rec["action"].append(action); rec["budget"].append(budget)
                    board = utils.step(board, action)
                    state = utils.board_to_state(board)
                    if utils.is_game_over(board, 5)[0]:
                        ok = len(rec["n"]) >= plies
                        break
            # ties inside the search would make n irreproducible; ties in the final pick only matter
            # when tau <= 0.01 (the uniform-over-best policy is still deterministic)
            if ok and tw.worst == 1:
                break
        else:
            raise AssertionError("no tie-free salt for the policy vectors")
        out[f"{tag}_salt"] = salt
        out[f"{tag}_init_temp"] = temp
        for k, v in rec.items():
            out[f"{tag}_{k}"] = np.asarray(v)
        print(f"policy {tag}: salt {salt}, tau {rec['tau'][0]:.4f} .. {rec['tau'][-1]:.5f}, budgets {rec['budget']}")
    # training mode, seeded
    config.init_temp = 1.2
    config.simulation_per_step, config.upper_simulation_per_step = 300, 400
    np.random.seed(99); random.seed(99)
    pl = Player(config, training=True, pv_fn=table_pv_fn(S, 9))
    state, action = pl.get_init_state(), None
    rec = dict(boards=[], n=[], policy=[], tau=[])
    for _ in range(12):
        policy, action = pl.get_action(state, last_action=action)
        rec["boards"].append(utils.state_to_board(state, S)); rec["n"].append(_root_arrays(pl, state, S)[0])
        rec["policy"].append(policy.reshape(-1).copy()); rec["tau"].append(pl.tau)
        board = utils.step(utils.state_to_board(state, S), action)
        state = utils.board_to_state(board)
        if utils.is_game_over(board, 5)[0]:
            break
    for k, v in rec.items():
        out[f"train_{k}"] = np.asarray(v)
    config.simulation_per_step, config.upper_simulation_per_step, config.init_temp = keep
    np.savez_compressed(os.path.join(OUT, "policy_vectors.npz"), **out)
    print("policy_vectors: train plies", len(rec["n"]))


def _mix_run(args):
    """One training-mode search of the real Player (300 sims, empty 11x11 board, table policy with
    zero value): root visit counts and the visit counts of every depth-1 node."""
    seed, salt, sims = args
    utils, config, Player = _ref()
    from oracle.mcts import table_pv_fn
    S = 11
    config.board_size = S
    config.simulation_per_step, config.upper_simulation_per_step = sims, sims + 100
    np.random.seed(seed); random.seed(seed)
    pl = Player(config, training=True, pv_fn=table_pv_fn(S, salt, zero_value=True))
    state = pl.get_init_state()
    pl.root_state = state
    for _ in range(sims):
        pl.MCTS_search(state, [state], None)
    root_n = _root_arrays(pl, state, S)[0]
    child_n = np.zeros((S * S, S * S), np.int16)
    empty = np.zeros((S, S), np.int8)
    for c in range(S * S):
        b = utils.step(empty.copy(), (c // S, c % S))
        st = utils.board_to_state(b)
        if st in pl.tree:
            child_n[c] = _root_arrays(pl, st, S)[0]
    return root_n, child_n


def make_mix_stats(runs=96, salt=9, sims=300):
    """Pins the Dirichlet mixing weights (player.py:247-253): with q == 0 everywhere the second visit of a
    depth-1 node picks argmax(0.9 p + 0.1 eta) and the post-ladder root picks follow
    (0.75 p + 0.25 eta) sqrt(sum_n + 1) / (1 + n)."""
    import multiprocessing as mp
    with mp.get_context("fork").Pool(8) as pool:
        res = pool.map(_mix_run, [(5000 + r, salt, sims) for r in range(runs)])
    np.savez_compressed(os.path.join(OUT, "mcts_mix_11.npz"), salt=salt, sims=sims,
                        root_n=np.stack([r[0] for r in res]), child_n=np.stack([r[1] for r in res]))
    print("mcts_mix_11:", runs, "runs")


def _game_run(args):
    seed, salt, sims, upper = args
    utils, config, Player = _ref()
    from oracle.mcts import table_pv_fn
    config.board_size = 11
    config.simulation_per_step, config.upper_simulation_per_step = sims, upper
    np.random.seed(seed); random.seed(seed)
    pl = Player(config, training=True, pv_fn=table_pv_fn(11, salt))
    rec = pl.run()
    value = rec[-1][-2]
    result = utils.DRAW if value == 0.0 else (utils.BLACK_WIN if len(rec) % 2 == 1 else utils.WHITE_WIN)   # main.py:86-93
    first = rec[1][2]                              # the first move played
    ent = [float(-(r[1][r[1] > 0] * np.log(r[1][r[1] > 0])).sum()) for r in rec[:8]]
    return len(rec), result, first[0] * 11 + first[1], ent


def make_games(games=400, salt=9, sims=60, upper=80, name="selfplay_games_11.npz"):
    """Whole self-play games of the real Player.run() (training mode, seeded per game): lengths, results
    (main.py:86-93), first moves and the policy entropy of the first 8 plies -- distributional pins."""
    import multiprocessing as mp
    utils = _ref()[0]
    with mp.get_context("fork").Pool(8) as pool:
        res = pool.map(_game_run, [(9000 + g, salt, sims, upper) for g in range(games)], chunksize=1)
    np.savez_compressed(os.path.join(OUT, name), salt=salt, sims=sims, upper=upper,
                        length=np.array([r[0] for r in res]), result=np.array([r[1] for r in res]),
                        first=np.array([r[2] for r in res]), entropy=np.array([r[3] for r in res], np.float32),
                        codes=np.array([utils.DRAW, utils.BLACK_WIN, utils.WHITE_WIN]))
    L = np.array([r[0] for r in res]); R = np.array([r[1] for r in res])
    print(name, games, "games; length mean", L.mean(), "min", L.min(), "max", L.max(),
          "black", (R == utils.BLACK_WIN).mean(), "white", (R == utils.WHITE_WIN).mean())


def make_replay(utils):
    data = pickle.load(open(f"{REF}/data_buffer/data6960.pkl", "rb"))
    lens = pickle.load(open(f"{REF}/data_buffer/data_len6960.pkl", "rb"))
    res = pickle.load(open(f"{REF}/data_buffer/result6960.pkl", "rb"))
    total = int(np.sum(lens))
    start0 = len(data) - total                    # game 0 may be front-truncated (utils.py:101-115)
    offs = np.concatenate([[0], np.cumsum(lens)]) + start0
    rng = np.random.default_rng(7)
    pick = np.sort(rng.choice(len(data), 1024, replace=False))
    games = [5, 123, 400]
    idx = list(pick)
    game_off = [0]
    gidx = []
    for g in games:
        gidx.extend(range(int(offs[g]), int(offs[g + 1])))
        game_off.append(len(gidx))

    def pack(ix):
        S = 11
        boards = np.stack([utils.state_to_board(data[i][0], S) for i in ix]).astype(np.int8)
        pol = np.stack([data[i][1] for i in ix]).astype(np.float32)
        la = np.array([data[i][2] if data[i][2] is not None else (-1, -1) for i in ix], np.int16)
        val = np.array([data[i][3] for i in ix], np.float32)
        wt = np.array([data[i][4] for i in ix], np.float32)
        st = np.array([data[i][0] for i in ix])
        return boards, pol, la, val, wt, st

    b, p, la, v, w, s = pack(idx)
    gb, gp, gla, gv, gw, gs = pack(gidx)
    np.savez_compressed(os.path.join(OUT, "replay_sample.npz"), boards=b, policy=p, last_action=la,
                        value=v, weight=w, states=s,
                        g_boards=gb, g_policy=gp, g_last_action=gla, g_value=gv, g_weight=gw,
                        g_states=gs, g_off=np.array(game_off), g_result=np.array([res[g] for g in games]),
                        logged_losses=np.array([2.155, 0.313, 2.145], np.float32))
    print("replay_sample: 1024 records +", len(gidx), "records of 3 games")


def make_ckpt():
    import hashlib
    from alphafive_b200 import ckpt
    w = ckpt.read_bundle(f"{REF}/ckpt")
    np.savez_compressed(os.path.join(OUT, "ckpt6960.npz"), **{k.replace("/", "__"): v for k, v in w.items()})
    # the shipped files themselves, to pin the bundle *writer*: the 1.7 KB index verbatim, the 3 MB data file by hash
    np.savez_compressed(os.path.join(OUT, "ckpt6960_files.npz"),
                        index=np.frombuffer(open(f"{REF}/ckpt/alphaFive-6960.index", "rb").read(), np.uint8),
                        data_sha256=hashlib.sha256(open(f"{REF}/ckpt/alphaFive-6960.data-00000-of-00001", "rb").read()).hexdigest(),
                        marker=open(f"{REF}/ckpt/checkpoint").read())
    print("ckpt6960:", len(w), "tensors", sum(v.size for v in w.values()), "params")


def make_weights(utils):
    L = np.arange(1, 65)
    ws = np.zeros((64, 64), np.float32)
    for l in L:
        ws[l - 1, :l] = utils.construct_weights(int(l), gamma=0.94)
    np.savez_compressed(os.path.join(OUT, "weights.npz"), w=ws)


def make_replay_stack(utils):
    """Drive the real RandomStack (utils.py:14-146) with games cut from the shipped replay buffer."""
    import io
    import contextlib
    data = pickle.load(open(f"{REF}/data_buffer/data6960.pkl", "rb"))
    lens = pickle.load(open(f"{REF}/data_buffer/data_len6960.pkl", "rb"))
    res = pickle.load(open(f"{REF}/data_buffer/result6960.pkl", "rb"))
    total = int(np.sum(lens))
    offs = np.concatenate([[0], np.cumsum(lens)]) + (len(data) - total)
    S = 11
    rng = np.random.default_rng(21)
    games = [int(g) for g in rng.choice(np.arange(1, len(lens)), 120, replace=False)]
    random.seed(1234)
    np.random.seed(4321)
    stack = utils.RandomStack(S, length=900)                # small: eviction happens many times
    accepted, n_data, black, white, first_len, n_games = [], [], [], [], [], []
    with contextlib.redirect_stdout(io.StringIO()):         # push prints statistics
        for g in games:
            game = data[int(offs[g]):int(offs[g + 1])]
            accepted.append(stack.push(list(game), int(res[g])))
            n_data.append(len(stack.data)); black.append(stack.black_win); white.append(stack.white_win)
            first_len.append(stack.data_len[0] if stack.data_len else 0); n_games.append(len(stack.data_len))
    boards, weights, values, policies = stack.get_data(256)
    la = np.array([r[2] if r[2] is not None else (-1, -1) for r in stack.data], np.int16)
    np.savez_compressed(
        os.path.join(OUT, "replay_stack.npz"), games=np.array(games), length=900,
        g_off=np.array([[int(offs[g]), int(offs[g + 1])] for g in games]), g_result=np.array([res[g] for g in games]),
        g_states=np.array([r[0] for g in games for r in data[int(offs[g]):int(offs[g + 1])]]),
        g_policy=np.stack([r[1] for g in games for r in data[int(offs[g]):int(offs[g + 1])]]).astype(np.float32),
        g_last=np.array([r[2] if r[2] is not None else (-1, -1) for g in games
                         for r in data[int(offs[g]):int(offs[g + 1])]], np.int16),
        g_value=np.array([r[3] for g in games for r in data[int(offs[g]):int(offs[g + 1])]], np.float32),
        g_weight=np.array([r[4] for g in games for r in data[int(offs[g]):int(offs[g + 1])]], np.float32),
        accepted=np.array(accepted), n_data=np.array(n_data), black=np.array(black), white=np.array(white),
        first_len=np.array(first_len), n_games=np.array(n_games),
        final_states=np.array([r[0] for r in stack.data]), final_last=la,
        final_data_len=np.array(stack.data_len), final_result=np.array(stack.result),
        batch_boards=boards, batch_weights=weights, batch_values=values, batch_policies=policies,
        seeds=np.array([1234, 4321]))
    print("replay_stack:", sum(accepted), "of", len(games), "games accepted,", len(stack.data), "records held")


def make_buffer_games():
    """Lengths and results of the 470 games in the shipped replay buffer (data_buffer/data_len6960.pkl,
    result6960.pkl): what utils.RandomStack (utils.py:64-116) had kept of the training-mode self-play around
    step 6960 -- a distributional pin for self-play with the trained net."""
    lens = np.array(pickle.load(open(os.path.join(REF, "data_buffer", "data_len6960.pkl"), "rb")), np.int32)
    res = np.array(pickle.load(open(os.path.join(REF, "data_buffer", "result6960.pkl"), "rb")), np.int8)
    assert len(lens) == len(res) == 470
    np.savez_compressed(os.path.join(OUT, "buffer_games_6960.npz"), lens=lens, results=res)
    print("buffer_games_6960.npz: %d games, mean length %.2f, black %d white %d" % (len(lens), lens.mean(), (res == 1).sum(), (res == -1).sum()))


def make_gui_game():
    """Decode tmp/five_6960.gif.  GUI.py appends one half-size screenshot per stone (frames 1..58) after the empty
    board, then five copies of the final position (GUI.py:176-181; self_play.py would append three, and writes
    tmp/five.gif).  pygame.surfarray.array3d is [x][y], so frame[row = x][col = y]; a stone (i, j) is a filled circle
    at x = (j + 1.5) * 36, y = (i + 1.5) * 36 (GUI.py draw_stone), black for the side that moved first."""
    from PIL import Image
    S, G = 11, 18
    im = Image.open(os.path.join(REF, "tmp", "five_6960.gif"))
    assert im.size == ((S + 2) * G, (S + 2) * G)
    boards = []
    for k in range(im.n_frames):
        im.seek(k)
        f = np.array(im.convert("RGB")).astype(int)
        b = np.zeros((S, S), np.int8)
        for i in range(S):
            for j in range(S):
                r, c = int((j + 1.5) * G), int((i + 1.5) * G)
                px = f[r - 2:r + 3, c - 2:c + 3].reshape(-1, 3).mean(0)
                b[i, j] = 1 if px.max() < 60 else (-1 if px.min() > 200 else 0)
        boards.append(b)
    assert not boards[0].any()
    moves = []
    for k in range(1, len(boards)):
        d = np.argwhere(boards[k] != boards[k - 1])
        if len(d) == 0:
            continue
        assert len(d) == 1, (k, d)
        i, j = (int(v) for v in d[0])
        assert boards[k - 1][i, j] == 0 and boards[k][i, j] == (1 if len(moves) % 2 == 0 else -1)
        moves.append((i, j))
    assert len(moves) == 58 and len(boards) == 1 + 58 + 5
    # consistency with utils: the game is over exactly after the last move, and the mover won
    utils = _ref()[0]
    board = np.zeros((S, S), np.int8)
    for t, mv in enumerate(moves):
        board = utils.step(board, mv)
        over, v = utils.is_game_over(board, 5)
        assert over == (t == len(moves) - 1)
    assert v == -1.0
    np.savez_compressed(os.path.join(OUT, "gui_game_6960.npz"), moves=np.array(moves, np.int8), ai_first=np.int8(1),
                        sims=np.int32(542), upper=np.int32(642), final_board=boards[-1])
    print("gui_game_6960.npz: 58 plies, the AI (black) moved first, the human (white) completed a five on ply 58")


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--gui-only" in sys.argv:
        make_gui_game()
        return
    if "--buffer-games-only" in sys.argv:
        make_buffer_games()
        return
    if "--ckpt-only" in sys.argv:
        make_ckpt()
        return
    if "--policy-only" in sys.argv:
        make_policy_vectors(*_ref())
        return
    if "--mix-only" in sys.argv:
        make_mix_stats()
        return
    if "--games-only" in sys.argv:
        make_games()
        return
    if "--games300-only" in sys.argv:
        make_games(games=200, sims=300, upper=380, name="selfplay_games_11_s300.npz")
        return
    if "--replay-stack-only" in sys.argv:
        make_replay_stack(_ref()[0])
        return
    utils, config, Player = _ref()
    if "--mcts-only" not in sys.argv:
        make_rules(11, utils, 3000, 11)
        make_rules(15, utils, 1500, 15)
        make_weights(utils)
        make_replay(utils)
        make_replay_stack(utils)
        make_ckpt()
    rs = np.load(os.path.join(OUT, "replay_sample.npz"))
    roots11 = [(np.zeros((11, 11), np.int8), None)]
    for t in (40, 333, 800):
        la = tuple(int(x) for x in rs["last_action"][t])
        roots11.append((rs["boards"][t].copy(), la if la[0] >= 0 else None))
    make_mcts_kat(11, utils, config, Player, roots11, [1, 2, 10, 100, 500], salt=3)
    rng = np.random.default_rng(5)
    from oracle import rules
    mid15 = rules.random_board(rng, 15, 0.12)
    while rules.terminal(mid15)[0]:
        mid15 = rules.random_board(rng, 15, 0.12)
    make_mcts_kat(15, utils, config, Player, [(np.zeros((15, 15), np.int8), None), (mid15, (7, 7))],
                  [1, 10, 120], salt=4)
    make_mcts_game(11, utils, config, Player, sims=120, upper=135, plies=14, salt=5)
    make_mcts_game(15, utils, config, Player, sims=60, upper=68, plies=8, salt=6)
    make_mcts_train(utils, config, Player, salt=9)
    make_policy_vectors(utils, config, Player)
    make_mix_stats()
    make_games()
    make_games(games=200, sims=300, upper=380, name="selfplay_games_11_s300.npz")
    make_gui_game()
    make_buffer_games()


if __name__ == "__main__":
    main()
