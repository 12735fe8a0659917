"""oracle/mcts.py against root statistics of the real reference Player (golden)."""
import numpy as np
import pytest

from conftest import golden
from oracle import mcts, rules


def _cfg(S, sims, upper):
    return mcts.SearchConfig(board_size=S, simulation_per_step=sims, upper_simulation_per_step=upper)


@pytest.mark.parametrize("S", [11, 15])
def test_deterministic_search_known_answers(S):
    g = golden(f"mcts_kat_{S}.npz")
    pv = mcts.table_pv_fn(S, int(g["salt"]))
    for t in range(len(g["k"])):
        k = int(g["k"][t])
        pl = mcts.OraclePlayer(_cfg(S, k, k + 100), training=False, pv_fn=pv)
        board = g["root_boards"][t]
        la = tuple(int(v) for v in g["root_last"][t])
        _, action = pl.get_action(board, la if la[0] >= 0 else None)
        n, w, p, sum_n = pl.root_stats(board)
        assert (n == g["n"][t]).all()
        assert sum_n == int(g["sum_n"][t])
        assert (w == g["w"][t]).all()                      # same f32 accumulation order
        assert (p == g["p"][t]).all()
        assert len(pl.table) == int(g["nkeys"][t])
        if g["action"][t][0] >= 0:
            assert action == tuple(int(v) for v in g["action"][t])
        lo, hi = int(g["key_off"][t]), int(g["key_off"][t + 1])
        want = {g["key_boards"][i].tobytes(): int(g["key_sum_n"][i]) for i in range(lo, hi)}
        got = {k_: v.sum_n for k_, v in pl.table.items()}
        assert got == want


@pytest.mark.parametrize("S", [11, 15])
def test_deterministic_game_with_tree_reuse(S):
    g = golden(f"mcts_game_{S}.npz")
    pl = mcts.OraclePlayer(_cfg(S, int(g["sims"]), int(g["upper"])), training=False,
                           pv_fn=mcts.table_pv_fn(S, int(g["salt"])))
    board, last = np.zeros((S, S), np.int8), None
    for t in range(len(g["action"])):
        assert (board == g["boards"][t]).all()
        assert pl.search_budget(pl.key_of(board)) == int(g["budget"][t])
        _, action = pl.get_action(board, last)
        n, w, _, sum_n = pl.root_stats(board)
        assert (n == g["n"][t]).all() and sum_n == int(g["sum_n"][t])
        assert (w == g["w"][t]).all()
        assert action == tuple(int(v) for v in g["action"][t])
        assert len(pl.table) == int(g["nkeys"][t])
        board, last = rules.play(board, action), action


def test_training_mode_distribution_matches_reference():
    """Forced-visit ladder + Dirichlet mixing: compare summary statistics of 64 seeded
    reference searches with 64 oracle searches (different RNG streams)."""
    g = golden("mcts_train_11.npz")
    ref_n = g["n"]
    S, sims = 11, int(g["sims"])
    rng = np.random.default_rng(99)
    ns, depth = [], []
    for _ in range(32):
        pl = mcts.OraclePlayer(_cfg(S, sims, sims + 100), training=True,
                               pv_fn=mcts.table_pv_fn(S, int(g["salt"])), rng=rng)
        board = np.zeros((S, S), np.int8)
        pl.root_key = pl.key_of(board)
        for _ in range(sims):
            pl.simulate(board, None)
        ns.append(pl.root_stats(board)[0])
        depth.append(pl.stat_selects / sims)
    ns = np.stack(ns)
    assert ns.min() >= 2 and ref_n.min() >= 2              # sims >= 2A+1 => every child twice
    assert (ns.sum(1) == sims - 1).all()
    assert abs(ns.max(1).mean() - ref_n.max(1).mean()) < 1.0
    assert abs(np.sort(ns, 1).mean(0) - np.sort(ref_n, 1).mean(0))[:-2].max() < 0.4
    assert abs(np.mean(depth) - g["depth"].mean()) < 0.08


def test_run_produces_reference_record_format():
    cfg = _cfg(11, 30, 40)
    pl = mcts.OraclePlayer(cfg, training=True, pv_fn=mcts.table_pv_fn(11, 1), rng=np.random.default_rng(3))
    rec = pl.run()
    L = len(rec)
    assert 9 <= L <= 121
    s0, pol0, la0, v_last, w_last = rec[0][0], rec[0][1], rec[0][2], rec[-1][3], rec[-1][4]
    assert s0 == "l/" * 11 and la0 is None and pol0.shape == (11, 11) and pol0.dtype == np.float32
    assert v_last in (1.0, 0.0, -0.0)
    vals = [r[3] for r in rec]
    assert all(vals[i + 1] == -vals[i] for i in range(L - 1))
    np.testing.assert_allclose([r[4] for r in rec], rules.ply_weights(L, cfg.gamma), atol=1e-6)
    assert mcts.game_result(rec) in (1, -1, 0)


# --------------------------------------------------------------------------- #
# calc_policy (player.py:84-126): vectors of the real Player over consecutive moves
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("tag", ["det_hi", "det_lo"])
def test_calc_policy_vectors_deterministic_search(tag):
    """training=False, random_a=True: deterministic search, soft policy.  The oracle player fed the golden
    roots reproduces n, the policy (bit-exact: same numpy ops) and tau (decays before use, player.py:108-111),
    including the tau <= 0.01 branch (det_lo crosses it at the 7th call)."""
    g = golden("policy_vectors.npz")
    S = 11
    cfg = _cfg(S, 120, 135)
    cfg.init_temp = float(g[f"{tag}_init_temp"])
    pl = mcts.OraclePlayer(cfg, training=False, pv_fn=mcts.table_pv_fn(S, int(g[f"{tag}_salt"])))
    taus = g[f"{tag}_tau"]
    assert (taus <= 0.01).any() == (tag == "det_lo")
    tau = cfg.init_temp
    for t in range(len(taus)):
        board = g[f"{tag}_boards"][t]
        la = tuple(int(v) for v in g[f"{tag}_last"][t])
        assert pl.search_budget(pl.key_of(board)) == int(g[f"{tag}_budget"][t])
        pl.root_key = pl.key_of(board)
        for _ in range(pl.search_budget(pl.root_key)):
            pl.simulate(board, la if la[0] >= 0 else None)
        n = pl.root_stats(board)[0]
        assert (n == g[f"{tag}_n"][t]).all()
        policy, action = pl.move_policy(board, random_a=True)
        tau *= cfg.tau_decay_rate_r
        assert pl.tau == tau == float(taus[t])
        want = g[f"{tag}_policy"][t]
        assert (policy.reshape(-1) == want).all(), np.abs(policy.reshape(-1) - want).max()
        if tau <= 0.01:                                   # uniform over the most-visited cells, plays one of them
            top = np.flatnonzero(want > 0)
            assert (n[top] == n.max()).all() and np.allclose(want[top], 1.0 / len(top))
            assert int(g[f"{tag}_action"][t][0]) * S + int(g[f"{tag}_action"][t][1]) in top


def test_calc_policy_vectors_training_counts():
    """(n, tau) -> policy on the noisy visit counts of a seeded training-mode game of the real Player."""
    g = golden("policy_vectors.npz")
    tau = 1.2
    for t in range(len(g["train_tau"])):
        tau *= 0.94
        assert tau == float(g["train_tau"][t])
        legal = g["train_boards"][t].reshape(-1) == 0
        pv = mcts.soft_policy(g["train_n"][t][legal], tau)
        want = g["train_policy"][t]
        assert (pv == want[legal]).all() and (want[~legal] == 0).all()
        assert abs(float(want.sum()) - 1) < 1e-5


def test_mix_weights_distribution_and_sensitivity():
    """Dirichlet mixing weights 0.25 (root) / 0.10 (elsewhere), player.py:247-253: prior-rank distribution
    of the visited cells (tests/mixstats.py) -- the oracle matches the reference runs, and the same test
    rejects an oracle with the two weights swapped (so the GPU test built on it has teeth)."""
    from scipy import stats
    import mixstats
    g = golden("mcts_mix_11.npz")
    salt, sims = int(g["salt"]), int(g["sims"])
    rr, rc = mixstats.prior_ranks(salt)
    ref_root, ref_child = mixstats.rank_samples(g["root_n"], g["child_n"], rr, rc)
    assert (g["root_n"] >= 2).all() and (g["root_n"].sum(1) == sims - 1).all()
    ok_root, ok_child = mixstats.rank_samples(*mixstats.oracle_counts(16, salt, sims, 1), rr, rc)
    sw_root, sw_child = mixstats.rank_samples(*mixstats.oracle_counts(16, salt, sims, 2, noise_mix=(0.1, 0.25)), rr, rc)
    assert stats.ks_2samp(ref_root, ok_root).statistic < mixstats.KS_ROOT
    assert stats.ks_2samp(ref_child, ok_child).statistic < mixstats.KS_CHILD
    assert stats.ks_2samp(ref_root, sw_root).statistic > 3 * mixstats.KS_ROOT
    assert stats.ks_2samp(ref_child, sw_child).statistic > 3 * mixstats.KS_CHILD
