"""Shared by the CPU (oracle) and GPU (device) distribution tests of the Dirichlet mixing weights
(player.py:247-253).  Scenario: training-mode search from the empty 11x11 board under the table policy
with zero value (q stays 0), ``sims`` >= 2A simulations.  Then

* every depth-1 node is selected from ``n_root[c] - 1`` times with scores (0.9 p + 0.1 eta) * sqrt(.)/(1 + n);
  which grandchildren get the visits depends on the non-root mix weight;
* after the forced-visit ladder (every root child twice, player.py:264-276) the root's remaining picks follow
  (0.75 p + 0.25 eta) sqrt(sum_n + 1) / (1 + n); which children get a third visit depends on the root weight.

The statistic is the *prior rank* (0 = largest prior) of the cells that received those visits: more noise
weight pushes visits towards low-prior cells.
"""
import numpy as np

from oracle import mcts, rules

S = 11
C = S * S
# two-sample KS bounds on the rank distributions (swapped weights give 0.27 / 0.36, see test_oracle_mcts.py)
KS_ROOT = 0.06
KS_CHILD = 0.03


def prior_ranks(salt):
    """rank_root[c], rank_child[c][cell] (-1 for the occupied cell) under table_pv_fn(S, salt)."""
    pv = mcts.table_pv_fn(S, salt, zero_value=True)
    empty = np.zeros((S, S), np.int8)
    p_root = pv(rules.input_planes(empty, None)[None])[0][0]
    rank_root = np.empty(C, np.int64)
    rank_root[np.argsort(-p_root, kind="stable")] = np.arange(C)
    xs = np.stack([rules.input_planes(rules.play(empty, (c // S, c % S)), (c // S, c % S)) for c in range(C)])
    p_child = pv(xs)[0]
    rank_child = np.full((C, C), -1, np.int64)
    for c in range(C):
        legal = np.array([k for k in range(C) if k != c])
        order = legal[np.argsort(-p_child[c][legal], kind="stable")]
        rank_child[c][order] = np.arange(C - 1)
    return rank_root, rank_child


def rank_samples(root_n, child_n, rank_root, rank_child):
    """root_n int[R, C], child_n int[R, C, C] -> (root ranks of the post-ladder visits,
    child ranks of all grandchild visits), each as a flat sample array."""
    root_n, child_n = np.asarray(root_n, np.int64), np.asarray(child_n, np.int64)
    extra = np.clip(root_n - 2, 0, None)
    root_s = np.repeat(np.broadcast_to(rank_root, root_n.shape).reshape(-1), extra.reshape(-1))
    rc = np.broadcast_to(rank_child, child_n.shape).reshape(-1)
    cn = child_n.reshape(-1)
    keep = rc >= 0
    child_s = np.repeat(rc[keep], cn[keep])
    return root_s, child_s


def oracle_counts(runs, salt, sims, seed, noise_mix=None):
    """The same scenario through oracle.mcts.OraclePlayer (optionally with other mix weights)."""
    rng = np.random.default_rng(seed)
    cfg = mcts.SearchConfig(board_size=S, simulation_per_step=sims, upper_simulation_per_step=sims + 100)
    empty = np.zeros((S, S), np.int8)
    root_n = np.zeros((runs, C), np.int64)
    child_n = np.zeros((runs, C, C), np.int64)
    for r in range(runs):
        pl = mcts.OraclePlayer(cfg, training=True, pv_fn=mcts.table_pv_fn(S, salt, zero_value=True), rng=rng)
        if noise_mix is not None:
            pl.noise_mix = noise_mix
        pl.root_key = pl.key_of(empty)
        for _ in range(sims):
            pl.simulate(empty, None)
        root_n[r] = pl.root_stats(empty)[0]
        for c in range(C):
            b = rules.play(empty, (c // S, c % S))
            if pl.key_of(b) in pl.table:
                child_n[r, c] = pl.root_stats(b)[0]
    return root_n, child_n
