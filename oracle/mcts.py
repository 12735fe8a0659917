"""Oracle (test infrastructure): the reference's transposition-table MCTS, restated.

Follows ``genData/player.py`` of the reference.  The search graph is a table keyed
by position (not a tree): different move orders reaching one position share one
node (player.py:28-29).  Nodes are stored as flat arrays over the legal cells in
row-major order (the order utils.py:238-245 gives the reference's per-node dict),
and the PUCT score reproduces the reference's dtype chain as numpy 2.x evaluates
it (player.py:261):

    t     = float32(c_puct) * p                      # python float * np.float32 -> f32
    u     = float64(t) * sqrt(float64(sum_n + 1)) / float64(1 + n)
    score = float32( float64(q) + u )                # stored into an f32 array

In training mode ``p`` is first mixed with a fresh Dirichlet draw at *every* node
visit, ``0.75 p + 0.25 eta`` at the root and ``0.9 p + 0.1 eta`` elsewhere, in
float64 (player.py:240-253).

Stochastic choices (Dirichlet, forced-visit picks, tie breaks, move sampling) use
a ``numpy.random.Generator`` handed in by the caller; the reference uses the
unseeded legacy global streams, so parity for those parts is distributional.  The
deterministic parts are pinned by ``tests/golden/mcts_kat_*.npz`` (root statistics
of the real reference ``Player`` under a tie-free table ``pv_fn``).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import rules


@dataclass
class Node:
    cells: np.ndarray                 # int   [A]  flat legal cells, row-major
    p: np.ndarray                     # f32   [A]  priors (frozen at first arrival)
    n: np.ndarray = field(default=None)   # int64 [A]  edge visit counts
    w: np.ndarray = field(default=None)   # f32   [A]  total action value
    q: np.ndarray = field(default=None)   # f32   [A]  w / n
    sum_n: int = 0                    # number of selections made *from* this node

    def __post_init__(self):
        a = self.cells.shape[0]
        self.n = np.zeros(a, np.int64)
        self.w = np.zeros(a, np.float32)
        self.q = np.zeros(a, np.float32)


class SearchConfig:
    """The attributes ``Player`` reads from the reference's ``config`` module
    (config.py:2-19), with the same names."""
    board_size = 11
    goal = 5
    simulation_per_step = 542
    upper_simulation_per_step = 642
    tau_decay_rate = 0.94
    tau_decay_rate_r = 0.9
    c_puct = 5.0
    dirichlet_alpha = 0.3
    gamma = 0.94
    init_temp = 1.2
    max_processes = 5

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


class OraclePlayer:
    """Restatement of ``genData.player.Player`` (player.py:23-284), pv_fn mode."""

    def __init__(self, cfg, training=True, pv_fn=None, rng=None):
        assert pv_fn is not None
        self.config = cfg
        self.training = training
        self.pv_fn = pv_fn
        self.rng = rng if rng is not None else np.random.default_rng(0)
        self.table: dict[bytes, Node] = {}
        self.root_key = None
        self.tau = cfg.init_temp
        # counters for the measurement formulas of SURVEY 8(d)
        self.stat_sims = self.stat_leaf_evals = self.stat_selects = self.stat_legal = 0

    # -- bookkeeping (player.py:37-51) ---------------------------------------
    def reset(self):
        self.table = {}
        self.root_key = None
        self.tau = self.config.init_temp

    @staticmethod
    def key_of(board: np.ndarray) -> bytes:
        return np.ascontiguousarray(board, np.int8).tobytes()

    # -- one move (player.py:128-147) -----------------------------------------
    def search_budget(self, key) -> int:
        node = self.table.get(key)
        if node is None:
            return self.config.simulation_per_step
        return min(self.config.simulation_per_step,
                   self.config.upper_simulation_per_step - node.sum_n)

    def get_action(self, board: np.ndarray, last_action=None, random_a=False):
        """``board`` is the position with the side to move as +1.  Returns
        ``(policy f32[S,S] | None, (i, j))`` like player.py:128-147."""
        board = np.asarray(board, np.int8)
        self.root_key = self.key_of(board)
        for _ in range(self.search_budget(self.root_key)):
            self.simulate(board, last_action)
        return self.move_policy(board, random_a)

    # -- one simulation (player.py:204-228 + 166-184) --------------------------
    def simulate(self, board: np.ndarray, last_action):
        S = self.config.board_size
        path = []                                   # (node, edge index) pairs
        self.stat_sims += 1
        while True:
            over, v = rules.terminal(board, self.config.goal)      # before lookup
            if over:
                break
            key = self.key_of(board)
            node = self.table.get(key)
            if node is None:
                v = self.expand(key, board, last_action)
                break
            e = self.select(node, key == self.root_key)
            path.append((node, e))
            cell = int(node.cells[e])
            last_action = (cell // S, cell % S)
            board = rules.play(board, last_action)
        # backup: the leaf itself gets nothing; sum_n is only touched by select
        for node, e in reversed(path):
            v = -v
            node.n[e] += 1
            node.w[e] = np.float32(node.w[e] + np.float32(v))
            node.q[e] = node.w[e] / np.float32(node.n[e])

    # -- leaf (player.py:186-202) ---------------------------------------------
    def expand(self, key, board, last_action):
        x = rules.input_planes(board, last_action)
        prob, value = self.pv_fn(x[None])
        prob, value = prob[0], value[0]
        self.stat_leaf_evals += 1
        cells = rules.legal_cells(board)
        total = np.float32(0.0)
        for c in cells:                              # sequential f32 sum, row-major
            total = np.float32(total + prob[c])
        # max(sum, 1e-5): a python float divisor is "weak" under NEP 50, so the
        # division is float32 / float32 either way
        denom = np.float32(1e-5) if np.float32(1e-5) > total else total
        pri = (prob[cells].astype(np.float32) / denom).astype(np.float32)
        self.table[key] = Node(cells=cells, p=pri)
        self.stat_legal += cells.shape[0]
        return value

    # -- PUCT selection (player.py:230-279) --------------------------------------
    def select(self, node: Node, is_root: bool) -> int:
        cfg = self.config
        node.sum_n += 1
        self.stat_selects += 1
        a = node.cells.shape[0]
        if self.training:
            eta = self.rng.dirichlet(cfg.dirichlet_alpha * np.ones(a))
            # mix weights 0.25 at the root / 0.10 elsewhere (player.py:249-253); ``noise_mix`` exists so the
            # distribution tests can show that they would notice swapped or wrong weights
            e_root, e_in = getattr(self, "noise_mix", (0.25, 0.1))
            if is_root:
                pm = (np.float32(1 - e_root) * node.p).astype(np.float64) + e_root * eta
            else:
                pm = (np.float32(1 - e_in) * node.p).astype(np.float64) + e_in * eta
            t = cfg.c_puct * pm                                   # float64
        else:
            t = (np.float32(cfg.c_puct) * node.p).astype(np.float64)
        u = t * np.sqrt(np.float64(node.sum_n + 1)) / (1 + node.n).astype(np.float64)
        score = (node.q.astype(np.float64) + u).astype(np.float32)
        if is_root and self.training:
            for k in (0, 1):                         # forced-visit ladder, root only
                idx = np.flatnonzero(node.n == k)
                if idx.size:
                    return int(self.rng.choice(idx))
        best = np.flatnonzero(score == score.max())
        return int(best[0] if best.size == 1 else self.rng.choice(best))

    # -- visit counts -> move (player.py:84-126) ---------------------------------
    def move_policy(self, board, random_a):
        cfg = self.config
        S = cfg.board_size
        node = self.table[self.key_of(board)]
        visits = node.n.astype(np.float32)
        top = np.flatnonzero(node.n == node.n.max())
        best = int(node.cells[top[0] if top.size == 1 else self.rng.choice(top)])
        best_action = (best // S, best % S)
        if not self.training and not random_a:
            return None, best_action
        self.tau *= cfg.tau_decay_rate_r if random_a else cfg.tau_decay_rate
        policy = np.zeros(S * S, np.float32)
        pv = soft_policy(node.n, self.tau)
        policy[node.cells] = pv
        if self.tau <= 0.01:
            return policy.reshape(S, S), best_action
        pick = int(node.cells[self.rng.choice(pv.shape[0], p=pv.astype(np.float64) / pv.astype(np.float64).sum())])
        return policy.reshape(S, S), (pick // S, pick % S)

    # -- whole game (player.py:53-82) ------------------------------------------
    def run(self):
        cfg = self.config
        S = cfg.board_size
        board = np.zeros((S, S), np.int8)
        last_action, over, value, rec = None, False, 0.0, []
        while not over:
            policy, action = self.get_action(board, last_action)
            rec.append((rules.encode_state(board), policy, last_action))
            board = rules.play(board, action)
            over, value = rules.terminal(board, cfg.goal)
            last_action = action
        self.reset()
        turns = len(rec)
        if turns % 2 == 1:
            value = -value
        weights = rules.ply_weights(turns, cfg.gamma)
        out = []
        for i in range(turns):
            out.append((*rec[i], value, weights[i]))
            value = -value
        return out

    # -- helpers for the parity tests ------------------------------------------
    def root_stats(self, board):
        """Dense per-cell arrays (n, w, p, legal mask) and sum_n of ``board``'s node."""
        S2 = self.config.board_size ** 2
        node = self.table[self.key_of(board)]
        n = np.zeros(S2, np.int64); w = np.zeros(S2, np.float32); p = np.zeros(S2, np.float32)
        n[node.cells], w[node.cells], p[node.cells] = node.n, node.w, node.p
        return n, w, p, node.sum_n


def soft_policy(n_legal, tau) -> np.ndarray:
    """player.py:112-120 on the visit counts of the legal cells (row-major), with the
    temperature *after* its decay: tau <= 0.01 -> uniform over the most-visited cells;
    else float32 ``(n / max n) ** (1 / tau)`` normalised by its float32 sum."""
    n_legal = np.asarray(n_legal)
    if tau <= 0.01:
        top = n_legal == n_legal.max()
        return np.where(top, np.float32(1.0 / top.sum()), np.float32(0)).astype(np.float32)
    pv = n_legal.astype(np.float32)
    pv /= np.max(pv)
    pv = np.power(pv, 1 / tau)                      # f32 array ** python float -> f32 (NEP 50)
    pv /= np.sum(pv)
    return pv


def game_result(record) -> int:
    """Outcome label of a finished game as the data-generating worker computes it
    (main.py:86-93): draw if the last stored value is 0, else black wins iff the
    number of plies is odd."""
    if record[-1][-2] == 0.0:
        return rules.DRAW
    return rules.BLACK_WIN if len(record) % 2 == 1 else rules.WHITE_WIN


def table_pv_fn(size: int, salt: int = 0, zero_value: bool = False):
    """A deterministic, transcendental-free, tie-free ``pv_fn`` for known-answer
    tests.  The policy over the S*S cells is a board-dependent
    affine permutation of distinct integer weights divided by their (exact) integer
    sum -- a single correctly-rounded float32 division per cell, so it is bit-portable
    across numpy builds -- and the value is an integer in [-1000, 1000] / 1000
    (``zero_value``: always 0, so q stays 0 and selection depends on the priors and the
    exploration noise only).  Vectorised over the batch (integer arithmetic: the values
    do not depend on the batch size)."""
    C = size * size
    P = 127 if C <= 127 else 227 if C <= 227 else 401   # a prime >= C
    assert C <= P
    idx = np.arange(C, dtype=np.int64)

    def fn(x):
        x = np.asarray(x)
        B = x.shape[0]
        planes = (x.reshape(B, 3, C) > 0.5).astype(np.int64)
        code = planes[:, 0] * 1 + planes[:, 1] * 2 + planes[:, 2] * 4        # [B, C]
        h = np.full(B, 1469598103 + salt, np.int64)
        for c in range(C):
            h = (h * 1000003 + code[:, c] * 7919 + c) % 2147483647
        a = 1 + h % (P - 1)
        off = (h // 131) % P
        wts = 1 + ((a[:, None] * idx[None, :] + off[:, None]) % P) * 3 + (idx % 3)[None, :]   # pairwise distinct
        tot = wts.sum(1)
        prob = wts.astype(np.float32) / tot.astype(np.float32)[:, None]
        val = ((h // 7) % 2001 - 1000).astype(np.float32) / np.float32(1000)
        if zero_value:
            val = np.zeros(B, np.float32)
        return prob.astype(np.float32), val.astype(np.float32)

    return fn
