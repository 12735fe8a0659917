"""The stochastic / soft paths of the device search against vectors and distributions generated from the
real reference Player (oracle/make_golden.py): calc_policy (player.py:84-126), move sampling (:125),
Dirichlet noise and its mixing weights (:240-253) and whole self-play games (player.py:53-82, main.py:86-93)."""
import numpy as np
import pytest
import torch
from scipy import stats

from conftest import golden
from oracle import mcts as omcts, rules as orules

pytestmark = pytest.mark.gpu


def _engine(S, n_games, sims, upper, **kw):
    from alphafive_b200.engine import SearchEngine, make_config
    return SearchEngine(make_config(board_size=S, simulation_per_step=sims, upper_simulation_per_step=upper,
                                    n_games=n_games, **kw))


def _cells(last, S):
    last = np.asarray(last).astype(np.int64)
    return np.where(last[:, 0] >= 0, last[:, 0] * S + last[:, 1], -1).astype(np.int32)


@pytest.mark.parametrize("tag", ["det_hi", "det_lo"])
def test_calc_policy_vectors(cuda_lib, tag):
    """Player(training=False).get_action(random_a=True) for 12-14 consecutive moves: the device holds the
    reference's visit counts exactly, so finish_move must return the reference's policy (<= 2e-6: powf vs
    np.power) and the reference's tau exactly -- soft branch (det_hi) and across tau <= 0.01 (det_lo)."""
    g = golden("policy_vectors.npz")
    S = 11
    temp = float(g[f"{tag}_init_temp"])
    eng = _engine(S, 1, 120, 135, training=False, random_a=True, init_temp=temp)
    pv = omcts.table_pv_fn(S, int(g[f"{tag}_salt"]))
    taus = g[f"{tag}_tau"]
    crossed = False
    for t in range(len(taus)):
        eng.set_roots(g[f"{tag}_boards"][t][None], _cells(g[f"{tag}_last"][t][None], S))
        assert int(eng.sims_left().cpu()[0]) == int(g[f"{tag}_budget"][t])
        eng.run_search(pv_fn=pv)
        n = eng.root_stats()[0].cpu().numpy()[0]
        assert (n == g[f"{tag}_n"][t]).all()
        policy, action = eng.finish_move()
        policy, action = policy.cpu().numpy()[0], int(action.cpu()[0])
        want = g[f"{tag}_policy"][t]
        assert float(eng.tau().cpu()[0]) == float(taus[t])                 # exact (double, decays before use)
        assert np.abs(policy - want).max() <= 2e-6, (t, np.abs(policy - want).max())
        assert (policy[want == 0] == 0).all()
        assert want[action] > 0                                            # the sampled move has support
        if taus[t] <= 0.01:
            crossed = True
            assert n[action] == n.max()                                    # plays a most-visited cell (player.py:115)
    assert crossed == (tag == "det_lo")
    eng.close()


def test_soft_policy_on_training_counts(cuda_lib):
    """training=True over 8 consecutive moves of 64 games: the policy the device returns equals the reference
    formula (oracle.mcts.soft_policy, pinned bit-exactly by policy_vectors.npz train_*) on the device's own
    noisy visit counts; tau follows 1.2 * 0.94^k exactly."""
    S, N, sims = 11, 64, 260
    eng = _engine(S, N, sims, sims + 100, training=True, seed=3)
    pv = omcts.table_pv_fn(S, 9)
    boards = np.zeros((N, S, S), np.int8)
    last = np.full(N, -1, np.int32)
    tau = 1.2
    for ply in range(8):                                  # no game can end before ply 9
        eng.set_roots(boards, last)
        eng.run_search(pv_fn=pv)
        n = eng.root_stats()[0].cpu().numpy()
        policy, action = eng.finish_move()
        policy, action = policy.cpu().numpy(), action.cpu().numpy()
        tau *= 0.94
        assert (eng.tau().cpu().numpy() == tau).all()
        for j in range(N):
            legal = boards[j].reshape(-1) == 0
            want = np.zeros(S * S, np.float32)
            want[legal] = omcts.soft_policy(n[j][legal], tau)
            assert np.abs(policy[j] - want).max() <= 2e-6
            assert legal[action[j]] and policy[j][action[j]] > 0
            boards[j] = orules.play(boards[j], (action[j] // S, action[j] % S))
            assert orules.terminal_code(boards[j]) == 0
        last = action.astype(np.int32)
    eng.close()


def test_sampled_moves_follow_policy_chi2(cuda_lib):
    """np.random.choice(A, p=pi) (player.py:125): 8192 games search the same root deterministically, so all
    hold the same pi; their sampled moves must be a multinomial sample of it."""
    S, N = 11, 8192
    g = golden("policy_vectors.npz")
    board, last = g["det_hi_boards"][3], g["det_hi_last"][3]
    eng = _engine(S, N, 120, 135, training=False, random_a=True, seed=17)
    eng.set_roots(np.repeat(board[None], N, 0), np.repeat(_cells(last[None], S), N))
    eng.run_search(pv_fn=omcts.table_pv_fn(S, int(g["det_hi_salt"])))
    policy, action = eng.finish_move()
    policy, action = policy.cpu().numpy(), action.cpu().numpy()
    assert (policy == policy[0]).all()
    pi = policy[0].astype(np.float64)
    pi /= pi.sum()
    counts = np.bincount(action, minlength=S * S).astype(np.float64)
    assert (counts[pi == 0] == 0).all()
    big = pi * N >= 5                                                      # pool the sparse cells
    obs = np.append(counts[big], counts[~big].sum())
    exp = np.append(pi[big] * N, pi[~big].sum() * N)
    if exp[-1] == 0:
        obs, exp = obs[:-1], exp[:-1]
    assert len(obs) >= 10
    chi2, p = stats.chisquare(obs, exp)
    assert p > 1e-4, (chi2, p)
    eng.close()


@pytest.mark.parametrize("A", [121, 98, 225])
def test_dirichlet_marginals(cuda_lib, A):
    """The noise vector of one node visit x 10^5 games: every coordinate of Dirichlet(0.3 * 1_A)
    (np.random.dirichlet, player.py:240) is Beta(0.3, 0.3 (A - 1)); rows sum to 1; partial sums of k
    coordinates are Beta(0.3 k, 0.3 (A - k))."""
    from alphafive_b200.engine import dirichlet_sample
    n = 100_000
    eta = dirichlet_sample(123 + A, 0.3, A, n).cpu().numpy().astype(np.float64)
    assert np.abs(eta.sum(1) - 1).max() < 1e-5 and (eta >= 0).all()
    for c in (0, 31, 32, A // 2, A - 1):
        d, p = stats.kstest(eta[:, c], stats.beta(0.3, 0.3 * (A - 1)).cdf)
        assert p > 1e-4, (c, d, p)
    k = 40
    d, p = stats.kstest(eta[:, 5:5 + k].sum(1), stats.beta(0.3 * k, 0.3 * (A - k)).cdf)
    assert p > 1e-4, (d, p)
    assert abs(eta.mean() - 1.0 / A) < 1e-9 + 1e-3 / A
    # a wrong alpha would be seen: Beta(0.25, .) / Beta(0.35, .) marginals are rejected outright
    for bad in (0.25, 0.35):
        assert stats.kstest(eta[:, 0], stats.beta(bad, bad * (A - 1)).cdf).pvalue < 1e-6
    # independent streams per game
    assert len({r[:4].tobytes() for r in eta[:2000]}) == 2000


def test_mix_weights_match_reference(cuda_lib):
    """0.75 p + 0.25 eta at the root, 0.9 p + 0.1 eta elsewhere (player.py:247-253): the prior-rank
    distribution of the visited cells (tests/mixstats.py) against 96 searches of the real Player.  The CPU
    test test_mix_weights_distribution_and_sensitivity shows the same bounds reject swapped weights 5-10x over."""
    import mixstats
    g = golden("mcts_mix_11.npz")
    S, C = 11, 121
    salt, sims = int(g["salt"]), int(g["sims"])
    rr, rc = mixstats.prior_ranks(salt)
    ref_root, ref_child = mixstats.rank_samples(g["root_n"], g["child_n"], rr, rc)
    N = 1024
    eng = _engine(S, N, sims, sims + 100, training=True, seed=29)
    empty = np.zeros((N, S, S), np.int8)
    eng.set_roots(empty, np.full(N, -1, np.int32))
    eng.run_search(pv_fn=omcts.table_pv_fn(S, salt, zero_value=True))
    root_n = eng.root_stats()[0].cpu().numpy()
    assert root_n.min() >= 2 and (root_n.sum(1) == sims - 1).all()
    child_n = np.zeros((N, C, C), np.int64)
    for c in range(C):
        b = orules.play(np.zeros((S, S), np.int8), (c // S, c % S))
        n, _, _, s = eng.node_stats(np.repeat(b[None], N, 0))
        child_n[:, c] = n.cpu().numpy()
        assert (s.cpu().numpy() == root_n[:, c] - 1).all()      # every visit but the expanding one selected from it
    dev_root, dev_child = mixstats.rank_samples(root_n, child_n, rr, rc)
    d_root = stats.ks_2samp(ref_root, dev_root).statistic
    d_child = stats.ks_2samp(ref_child, dev_child).statistic
    assert d_root < mixstats.KS_ROOT and d_child < mixstats.KS_CHILD, (d_root, d_child)
    eng.close()


@pytest.mark.parametrize("fixture", ["selfplay_games_11.npz", "selfplay_games_11_s300.npz"])
def test_whole_game_distribution(cuda_lib, fixture):
    """Player.run() in training mode under the table pv_fn: 2048 device games against seeded games of the real
    Player (400 at 60 sims/move: the root never leaves the forced-visit ladder; 200 at 300 sims/move: PUCT with
    noise decides) -- two-sample KS on the game length, two-proportion z on the black win rate (main.py:86-93)
    and the draw rate, mean policy entropy of the first plies."""
    import os
    from conftest import GOLDEN
    from alphafive_b200.engine import parse_records
    if not os.path.exists(os.path.join(GOLDEN, fixture)):
        pytest.skip(f"{fixture} not generated")
    g = golden(fixture)
    S, N = 11, 2048
    sims, upper = int(g["sims"]), int(g["upper"])
    eng = _engine(S, N, sims, upper, training=True, auto_play=True, seed=41, record_capacity=N * 121)
    pv = omcts.table_pv_fn(S, int(g["salt"]))
    prob = torch.empty((N, S * S), device="cuda")
    value = torch.empty((N,), device="cuda")
    eng.step()
    first = {}
    for it in range(121 * (sims + 1) + 200):
        x = eng.planes().cpu().numpy()
        p, v = pv(x)
        prob.copy_(torch.from_numpy(p)); value.copy_(torch.from_numpy(v))
        eng.step(prob, value)
        if it % 512 == 511:
            buf, _ = eng.harvest()
            for r in parse_records(buf, S):
                if r["game_serial"] == 0:                                  # the first game of every slot: no length bias
                    first.setdefault(r["game_id"], []).append(r)
            if len(first) == N:
                break
    c = eng.counters()
    assert c["overflows"] == 0 and c["records_dropped"] == 0
    assert len(first) == N and all(len(v_) == v_[0]["game_len"] for v_ in first.values())
    length = np.array([v_[0]["game_len"] for v_ in first.values()])
    result = np.array([v_[0]["result"] for v_ in first.values()])
    ks = stats.ks_2samp(length, g["length"])
    assert ks.pvalue > 1e-3, (ks, length.mean(), g["length"].mean())

    def two_prop(k1, n1, k2, n2):
        p = (k1 + k2) / (n1 + n2)
        return (k1 / n1 - k2 / n2) / max(np.sqrt(p * (1 - p) * (1 / n1 + 1 / n2)), 1e-9)

    codes = g["codes"]                                                     # DRAW, BLACK_WIN, WHITE_WIN
    for code in codes[:2]:
        z = two_prop((result == code).sum(), N, (g["result"] == code).sum(), len(g["result"]))
        assert abs(z) < 3.5, (int(code), z)
    ent = np.zeros((N, 8))
    for i, v_ in enumerate(first.values()):
        v_.sort(key=lambda r: r["ply"])
        for t in range(8):
            pol = v_[t]["policy"].reshape(-1)
            ent[i, t] = -(pol[pol > 0] * np.log(pol[pol > 0])).sum()
    ref_ent = g["entropy"]
    se = np.sqrt(ent.var(0) / N + ref_ent.var(0) / len(ref_ent))
    assert (np.abs(ent.mean(0) - ref_ent.mean(0)) < 4.0 * se + 2e-3).all(), (ent.mean(0), ref_ent.mean(0))
    eng.close()


def test_self_play_with_the_trained_net_fills_the_buffer_like_the_reference(cuda_lib):
    """First-party distributional pin of training-mode self-play with the real network: the reference ships the
    replay buffer its five workers had filled around step 6960 (470 games after utils.RandomStack's short-game
    rejection, colour re-balancing duplicates and FIFO eviction, utils.py:64-116).  Device self-play with ckpt-6960
    at config.py's 542 / 642 simulations, pushed through the same bookkeeping, must leave a buffer of the same
    character: mean game length 25.5 (sd 8.5), 53 % black wins.  (The shipped games come from the checkpoints
    *before* 6960, so the match is statistical and the bounds are ~3 standard errors of a 470-game buffer plus that
    drift; measured: 24.5 +- 0.24 plies, 56 % black.)"""
    import random
    from alphafive_b200.engine import parse_records
    from alphafive_b200.net import DeviceNet
    from alphafive_b200.selfplay import SelfPlay
    g = golden("buffer_games_6960.npz")
    z = golden("ckpt6960.npz")
    N, sims, upper = 1024, 542, 642
    net = DeviceNet(11, N, {k.replace("__", "/"): z[k] for k in z.files})
    sp = SelfPlay(None, n_games=N, net=net, training=True, seed=11, board_size=11, simulation_per_step=sims,
                  upper_simulation_per_step=upper)
    lens, res = [], []
    while len(lens) < 1500:
        sp.run_passes(sims * 4)
        seen = set()
        for r in parse_records(sp.harvest()[0], 11):
            if (r["game_id"], r["game_serial"]) not in seen:
                seen.add((r["game_id"], r["game_serial"]))
                lens.append(int(r["game_len"])); res.append(int(r["result"]))
    assert sp.engine.counters()["overflows"] == 0
    means, blacks = [], []
    for seed in range(8):                                  # the bookkeeping of utils.py:64-116 on (length, result)
        rnd, L, R, bw, ww, total = random.Random(seed), [], [], 0, 0, 0
        for l, r in zip(lens, res):
            if rnd.random() <= -0.0682 * l + 1.364:
                continue
            L.append(l); R.append(r); total += l
            if r == 1:
                bw += 1
                if rnd.random() < (ww - bw) / (bw * 1.3):
                    L.append(l); R.append(r); total += l; bw += 1
            elif r == -1:
                ww += 1
                if rnd.random() < (bw - ww) / (ww * 1.02):
                    L.append(l); R.append(r); total += l; ww += 1
            while total > 12000:
                total -= L.pop(0)
                r0 = R.pop(0)
                bw -= r0 == 1; ww -= r0 == -1
        means.append(np.mean(L)); blacks.append(np.mean(np.array(R) == 1))
    ref_mean, ref_black = g["lens"].mean(), (g["results"] == 1).mean()
    assert abs(np.mean(means) - ref_mean) < 2.2, (np.mean(means), ref_mean)
    assert abs(np.mean(blacks) - ref_black) < 0.08, (np.mean(blacks), ref_black)
    assert 9 <= min(lens) and max(lens) <= 121 and (np.array(res) == 0).mean() < 0.02
    sp.engine.close(); net.close()
