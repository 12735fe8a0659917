"""``Player`` with the surface of the reference's ``genData.player.Player`` (player.py:23-284),
backed by the device search engine.

Same constructor, same methods, same return types: ``get_action(state, e, last_action,
random_a) -> (policy f32[S,S] | None, (i, j))``, ``run() -> [(state, policy, last_action,
value, weight)]``, ``reset``, ``get_init_state``, ``pruning_tree``, ``close``.  One Player is
one game slot of a ``SearchEngine`` (batch 1); many games at once go through
``alphafive_b200.selfplay``.  The leaf evaluator is whatever the caller injects, as in the
reference: a ``pv_fn`` callable (player.py:190-192) or a Pipe to a ``NetworkAPI`` server
(player.py:194-197); when ``pv_fn`` is the ``eval`` of this package's ``ResNet`` the
evaluation stays on the device.
"""
from __future__ import annotations

import gc
from collections.abc import Mapping

import numpy as np
import torch

from .. import rules
from .._lib import A5Error
from ..engine import SearchEngine, make_config


# -- host-side string <-> board conversion at the API boundary (utils.py:156-196) ---------
def board_to_state(board) -> str:
    parts = []
    for row in np.asarray(board).tolist():
        gap = 0
        for v in row:
            if v:
                if gap:
                    parts.append(chr(97 + gap))
                    gap = 0
                parts.append(str(v + 2))
            else:
                gap += 1
        parts.append((chr(97 + gap) if gap else "") + "/")
    return "".join(parts)


def state_to_board(state: str, board_size: int) -> np.ndarray:
    board = np.zeros((board_size, board_size), np.int8)
    for r, row in enumerate(state.split("/")[:board_size]):
        c = 0
        for ch in row:
            if ch.isalpha():
                c += ord(ch) - 97
            else:
                board[r, c] = int(ch) - 2
                c += 1
    return board


class Action(object):
    """Edge statistics as the reference holds them (player.py:9-14); read-only copy of the device node."""
    __slots__ = ("n", "w", "q", "p")

    def __init__(self, n=0, w=0.0, p=0.0):
        self.n, self.w, self.p = n, w, p
        self.q = w / n if n else 0


class State(object):
    """Node as the reference holds it (player.py:17-20): ``a`` maps every legal (i, j) to its Action."""
    __slots__ = ("a", "sum_n")

    def __init__(self, a, sum_n):
        self.a, self.sum_n = a, sum_n


class _TreeView(Mapping):
    """Read-only ``Player.tree`` (player.py:29): state string -> State, fetched from the device table on
    access (a5_engine_table_dump for the keys, a5_engine_node_stats for a node)."""

    def __init__(self, engine, S):
        self._eng, self._S = engine, S
        self._keys = None

    def _load(self):
        if self._keys is None:
            boards, sums = ([], []) if self._eng is None else self._eng.table_dump(0)
            self._keys = {board_to_state(b): int(s) for b, s in zip(boards, sums)}
        return self._keys

    def __len__(self):
        return len(self._load())

    def __iter__(self):
        return iter(self._load())

    def __contains__(self, state):
        return state in self._load()

    def __getitem__(self, state):
        if state not in self._load():
            raise KeyError(state)
        S = self._S
        board = state_to_board(state, S)
        n, w, p, s = [t.cpu().numpy()[0] for t in self._eng.node_stats(board[None])]
        a = {(c // S, c % S): Action(int(n[c]), np.float32(w[c]), np.float32(p[c]))
             for c in np.flatnonzero(board.reshape(-1) == 0)}
        return State(a, int(s))


class Player(object):
    def __init__(self, cfg=None, training=True, pipe=None, pv_fn=None, seed=0):
        assert pipe is not None or pv_fn is not None
        self.config = cfg
        self.training = training
        self.pipe = pipe
        self.pv_fn = pv_fn
        self.root_state = None
        self.goal = self.config.goal
        self.tau = self.config.init_temp
        self.job_done = False
        self._seed = seed
        self._engine = None
        self._budget = None
        self._clear = True
        from .._lib import NET_SMALL
        self.net_mode = NET_SMALL                    # compute path of an on-device pv_fn (None: the net's own)

    # ------------------------------------------------------------------ engine plumbing
    def _eng(self) -> SearchEngine:
        cfg = self.config
        budget = (cfg.simulation_per_step, cfg.upper_simulation_per_step)
        # the table holds the retained sub-tree (at most ~upper nodes: its root was selected from that often)
        # plus this move's expansions; sized from the budget in force and re-created (tree dropped, like a
        # reset) if the caller raises config.simulation_per_step past it (choose_best_player.py:25 mutates it)
        need = min(max(2 * budget[0], budget[0] + budget[1]) + 1024, 32768)
        if self._engine is None or self._engine.S != cfg.board_size or need > self._engine.cfg.node_capacity:
            if self._engine is not None:
                self._engine.close()
            self._engine = SearchEngine(make_config(cfg, n_games=1, training=self.training, seed=self._seed,
                                                    node_capacity=need))
            self._clear = True
            self._budget = None
        if budget != self._budget:                   # config is read lazily
            self._engine.set_budget(*budget)
            self._budget = budget
        return self._engine

    def _leaf_fn(self):
        if self.pv_fn is not None:
            owner = getattr(self.pv_fn, "__self__", None)
            if getattr(owner, "device_net", None) is not None and getattr(self.pv_fn, "__name__", "") == "eval":
                return owner.device_net, None        # evaluate on the device
            return None, self.pv_fn

        def via_pipe(x):                             # player.py:194-197
            out_p, out_v = [], []
            for xi in x:
                self.pipe.send([xi])
                while not self.pipe.poll():
                    pass
                p, v = self.pipe.recv()[0]
                out_p.append(p)
                out_v.append(v)
            return np.asarray(out_p, np.float32), np.asarray(out_v, np.float32)
        return None, via_pipe

    # ------------------------------------------------------------------ reference API
    def get_init_state(self):
        return (chr(ord("a") + self.config.board_size) + "/") * self.config.board_size

    def reset(self, search_tree=None):
        """player.py:48-51.  The table lives on the device: a caller-supplied dict of State objects cannot
        be adopted, so anything but ``None`` is an error rather than being silently ignored."""
        if search_tree is not None:
            raise NotImplementedError("Player.reset(search_tree=...): the search table is device-resident; "
                                      "only reset() / reset(None) is supported")
        self._clear = True
        self.root_state = None
        self.tau = self.config.init_temp

    @property
    def tree(self):
        """The transposition table as the reference exposes it (player.py:29): a read-only mapping
        state string -> State(a = {(i, j): Action(n, w, q, p)}, sum_n), read from the device on access."""
        if self._engine is None or self._clear:
            return _TreeView(None, self.config.board_size)
        return _TreeView(self._engine, self.config.board_size)

    def get_action(self, state: str, e: float = 0.25, last_action: tuple = None, random_a=False):
        eng = self._eng()
        S = self.config.board_size
        self.root_state = state
        board = state_to_board(state, S)
        last = -1 if last_action is None else int(last_action[0]) * S + int(last_action[1])
        eng.set_mode(self.training, random_a)
        eng.set_roots(board[None], np.array([last], np.int32), None,
                      np.array([1 if self._clear else 0], np.uint8))
        self._clear = False
        net, fn = self._leaf_fn()
        # one board, strictly sequential leaf evaluations: the one-kernel latency path of the device net
        eng.run_search(net=net, pv_fn=fn, check_every=8, net_mode=None if net is None else self.net_mode)
        try:
            eng.counters()                            # raises A5_ERR_CAPACITY if an expansion did not fit
        except A5Error as err:
            raise A5Error(f"Player.get_action: search table overflow ({err}); visit counts would diverge "
                          "from the reference") from None
        policy, action = eng.finish_move()
        a = int(action.cpu()[0])
        act = (a // S, a % S)
        if not self.training and not random_a:
            return None, act
        self.tau *= self.config.tau_decay_rate_r if random_a else self.config.tau_decay_rate
        return policy.cpu().numpy().reshape(S, S), act

    def run(self, e=0.25):
        """One self-play game; returns [(state, policy, last_action, value, weight)] (player.py:53-82)."""
        S = self.config.board_size
        state = self.get_init_state()
        data, value, last_action, over = [], 0.0, None, False
        while not over:
            policy, action = self.get_action(state, e, last_action)
            data.append((state, policy, last_action))
            board = torch.from_numpy(state_to_board(state, S)[None]).cuda()
            nxt = rules.step(board, torch.tensor([action[0] * S + action[1]], dtype=torch.int32, device="cuda"))
            code = int(rules.terminal(nxt, self.goal).cpu()[0])
            state = board_to_state(nxt[0].cpu().numpy())
            over = code != 0
            value = {0: 0.0, 1: 1.0, 2: -1.0, 3: 0.0}[code]
            last_action = action
        self.reset()
        turns = len(data)
        if turns % 2 == 1:
            value = -value
        weights = construct_weights(turns, gamma=self.config.gamma)
        out = []
        for i in range(turns):
            out.append((*data[i], value, weights[i]))
            value = -value
        return out

    def pruning_tree(self, board: np.ndarray, state: str = None):
        """The reference deletes the same-perspective ancestors of ``board`` here (player.py:149-164:
        keys != state whose +1 stones and -1 stones are subsets of the board's).  The device table is
        garbage-collected against the root at every get_action -- only positions containing all of the
        root's stones survive -- which removes a superset of those keys (tests/test_gpu_facade.py), so there
        is nothing left to delete."""
        return None

    def close(self):
        self.job_done = True
        if self._engine is not None:
            self._engine.close()
            self._engine = None
        gc.collect()


def construct_weights(length: int, gamma=0.95):
    """utils.py:286-296 (host side of Player.run; the device computes the same in emit_game)."""
    w = np.empty((int(length),), np.float32)
    w[length - 1] = 1.0
    for i in range(length - 2, -1, -1):
        w[i] = w[i + 1] * gamma
    return length * w / np.sum(w)
