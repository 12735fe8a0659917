"""Device search engine (C ABI) vs golden root statistics of the real reference Player,
vs the oracle, and self-play record invariants."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import mcts as omcts, rules as orules

pytestmark = pytest.mark.gpu


def _engine(S, n_games, sims, upper, **kw):
    from alphafive_b200.engine import SearchEngine, make_config
    cfg = make_config(board_size=S, simulation_per_step=sims, upper_simulation_per_step=upper,
                      n_games=n_games, **kw)
    return SearchEngine(cfg)


def _cells(last, S):
    last = np.asarray(last).astype(np.int64)
    return np.where(last[:, 0] >= 0, last[:, 0] * S + last[:, 1], -1).astype(np.int32)


@pytest.mark.parametrize("S", [11, 15])
def test_deterministic_known_answers(cuda_lib, S):
    """training=False search under the tie-free table pv_fn: root n / w / p / sum_n, the
    chosen move and the whole table (keys + sum_n) equal the reference's, exactly."""
    g = golden(f"mcts_kat_{S}.npz")
    pv = omcts.table_pv_fn(S, int(g["salt"]))
    ks = g["k"]
    for k in np.unique(ks):
        idx = np.flatnonzero(ks == k)
        eng = _engine(S, len(idx), int(k), int(k) + 100, training=False)
        eng.set_roots(g["root_boards"][idx], _cells(g["root_last"][idx], S))
        eng.run_search(pv_fn=pv)
        n, w, p, s = [t.cpu().numpy() for t in eng.root_stats()]
        _, action = eng.finish_move()
        action = action.cpu().numpy()
        for j, t in enumerate(idx):
            assert (n[j] == g["n"][t].reshape(-1)).all(), (S, k, j)
            assert s[j] == g["sum_n"][t]
            assert (w[j] == g["w"][t].reshape(-1)).all()
            assert (p[j] == g["p"][t].reshape(-1)).all()
            if g["action"][t][0] >= 0:
                assert action[j] == g["action"][t][0] * S + g["action"][t][1]
            boards, sums = eng.table_dump(j)
            lo, hi = int(g["key_off"][t]), int(g["key_off"][t + 1])
            want = {g["key_boards"][i].tobytes(): int(g["key_sum_n"][i]) for i in range(lo, hi)}
            got = {b.tobytes(): int(v) for b, v in zip(boards, sums)}
            assert got == want
        c = eng.counters()
        assert c["overflows"] == 0
        eng.close()


@pytest.mark.parametrize("S", [11, 15])
def test_deterministic_game_with_tree_reuse(cuda_lib, S):
    """get_action move after move on one engine: retained sub-tree statistics, the
    budget rule min(sims, upper - sum_n) and the moves equal the reference's."""
    g = golden(f"mcts_game_{S}.npz")
    sims, upper = int(g["sims"]), int(g["upper"])
    eng = _engine(S, 1, sims, upper, training=False)
    pv = omcts.table_pv_fn(S, int(g["salt"]))
    for t in range(len(g["action"])):
        eng.set_roots(g["boards"][t][None], _cells(g["last"][t][None], S))
        assert int(eng.sims_left().cpu()[0]) == int(g["budget"][t])
        eng.run_search(pv_fn=pv)
        n, w, _, s = [x.cpu().numpy() for x in eng.root_stats()]
        assert (n[0] == g["n"][t].reshape(-1)).all() and s[0] == g["sum_n"][t]
        assert (w[0] == g["w"][t].reshape(-1)).all()
        _, action = eng.finish_move()
        assert int(action.cpu()[0]) == g["action"][t][0] * S + g["action"][t][1]
    eng.close()


def test_matches_oracle_on_many_roots_with_device_net(cuda_lib):
    """N different mid-game roots searched in lock-step with the on-device net.  (i) The oracle search driven
    by the *same* device net (DeviceNet.eval as its pv_fn) must give identical visit counts on every root:
    the tree pass itself is exact.  (ii) With the fp32 oracle net instead, a root may differ only because the
    two nets round differently (<= 1e-4, tested in test_gpu_net.py) and flip a near-tie; that is rare."""
    from alphafive_b200.net import DeviceNet, glorot_init
    from oracle import net as onet
    S, sims = 11, 64
    g = golden("replay_sample.npz")
    boards, last = g["boards"][:48], g["last_action"][:48]
    w = glorot_init(S, 3)
    net = DeviceNet(S, 48, w)
    eng = _engine(S, 48, sims, sims + 100, training=False)
    eng.set_roots(boards, _cells(last, S))
    eng.run_search(net=net, check_every=4)
    n = eng.root_stats()[0].cpu().numpy()
    onet_ = onet.OracleNet(S, w)
    same = checked = 0
    for j in range(48):
        if not boards[j].any():
            continue          # the empty board: all-zero input + zero biases = exactly uniform prior, a 121-way tie (random pick)
        checked += 1
        la = tuple(int(v) for v in last[j])
        la = la if la[0] >= 0 else None
        cfg = omcts.SearchConfig(simulation_per_step=sims, upper_simulation_per_step=sims + 100)
        pl = omcts.OraclePlayer(cfg, training=False, pv_fn=net.eval)
        pl.get_action(boards[j], la)
        assert (pl.root_stats(boards[j])[0] == n[j]).all(), j
        pl = omcts.OraclePlayer(cfg, training=False, pv_fn=onet_.eval)
        pl.get_action(boards[j], la)
        same += int((pl.root_stats(boards[j])[0] == n[j]).all())
        assert n[j].sum() == sims - 1
    assert checked >= 46 and same >= checked - 3, (same, checked)


def test_training_mode_distribution(cuda_lib):
    """Forced-visit ladder and per-visit Dirichlet mixing: summary statistics against 64
    seeded searches of the reference (tests/golden/mcts_train_11.npz)."""
    g = golden("mcts_train_11.npz")
    ref = g["n"]
    S, sims, N = 11, int(g["sims"]), 256
    eng = _engine(S, N, sims, sims + 100, training=True, seed=11)
    eng.set_roots(np.zeros((N, S, S), np.int8), np.full(N, -1, np.int32))
    eng.run_search(pv_fn=omcts.table_pv_fn(S, int(g["salt"])))
    n = eng.root_stats()[0].cpu().numpy()
    c = eng.counters()
    assert n.min() >= 2 and (n.sum(1) == sims - 1).all()
    assert abs(n.max(1).mean() - ref.max(1).mean()) < 0.6
    assert np.abs(np.sort(n, 1).mean(0) - np.sort(ref, 1).mean(0))[:-2].max() < 0.3
    depth = c["selects"] / c["sims"]
    assert abs(depth - g["depth"].mean()) < 0.05, depth
    assert len({r.tobytes() for r in n}) > N // 2            # independent Philox streams per game
    eng.close()


def test_forced_visit_picks_are_uniform(cuda_lib):
    """Root, training: while unvisited children exist one is drawn uniformly (player.py:264-271)."""
    S, N = 11, 4096
    eng = _engine(S, N, 3, 103, training=True, seed=5)

    def uniform_pv(x):
        B = x.shape[0]
        return np.full((B, S * S), 1.0 / (S * S), np.float32), np.zeros(B, np.float32)

    eng.set_roots(np.zeros((N, S, S), np.int8), np.full(N, -1, np.int32))
    eng.run_search(pv_fn=uniform_pv)
    n = eng.root_stats()[0].cpu().numpy()
    assert (n.sum(1) == 2).all() and n.max() == 1           # forced ladder: two distinct unvisited cells
    freq = n.sum(0) / (2 * N)
    assert abs(freq - 1 / 121).max() < 0.004                # uniform picks among the unvisited
    eng.close()


@pytest.mark.parametrize("S,N,sims,want", [(11, 64, 24, 40), (15, 32, 16, 16)])
def test_self_play_records(cuda_lib, S, N, sims, want):
    """auto_play: whole games on the device (Player.run + gen_data): record invariants of
    player.py:53-82 / main.py:86-93 / utils.py:286-296, at both board sizes of BASELINE.json."""
    from alphafive_b200.engine import parse_records
    from alphafive_b200.net import DeviceNet
    net = DeviceNet(S, N)
    eng = _engine(S, N, sims, sims + 10, training=True, auto_play=True, seed=1)
    prob = torch.empty((N, S * S), device="cuda")
    value = torch.empty((N,), device="cuda")
    eng.step()
    recs, games = [], 0
    for it in range(40000):
        net.forward_raw(eng.planes_ptr, N, prob, value)
        eng.step(prob, value)
        if it % 500 == 499:
            buf, g_ = eng.harvest()
            recs += parse_records(buf, S)
            games += g_
            if games >= want:
                break
    assert games >= want
    c = eng.counters()
    assert c["overflows"] == 0 and c["records_dropped"] == 0 and c["games"] >= games
    by_game = {}
    for r in recs:
        by_game.setdefault((r["game_id"], r["game_serial"]), []).append(r)
    assert len(by_game) == games
    lens = []
    for key, plies in by_game.items():
        plies.sort(key=lambda r: r["ply"])
        L = plies[0]["game_len"]
        lens.append(L)
        assert [r["ply"] for r in plies] == list(range(L)) and 9 <= L <= S * S
        assert not plies[0]["board"].any() and plies[0]["last_action"] == -1
        for t in range(L - 1):
            a = plies[t + 1]["last_action"]
            assert plies[t]["board"].reshape(-1)[a] == 0
            assert (orules.play(plies[t]["board"], (a // S, a % S)) == plies[t + 1]["board"]).all()
            assert orules.terminal_code(plies[t + 1]["board"]) == 0
        for r in plies:
            pol = r["policy"].reshape(-1)
            assert abs(pol.sum() - 1) < 1e-4 and (pol[r["board"].reshape(-1) != 0] == 0).all()
        final_moves = [orules.terminal(orules.play(plies[-1]["board"], a)) for a in orules.legal_actions(plies[-1]["board"])]
        vals = np.array([r["value"] for r in plies])
        if plies[0]["result"] == 0:
            assert (vals == 0).all() and (True, 0.0) in final_moves
        else:
            assert vals[-1] == 1.0 and (vals[1:] == -vals[:-1]).all() and (True, -1.0) in final_moves
            assert plies[0]["result"] == (1 if L % 2 == 1 else -1)
        np.testing.assert_allclose([r["weight"] for r in plies], orules.ply_weights(L, 0.94), atol=1e-6)
    assert np.mean(lens) > 12
    eng.close()


def test_record_arena_overflow_is_loud_and_never_splits_a_game(cuda_lib):
    """A record arena that is too small: harvest raises (A5_ERR_CAPACITY), counters()['records_dropped'] counts
    the lost plies, and what *is* returned are whole games with contiguous plies (emit_game reserves a
    game's plies atomically -- no roll-back that could hand two warps overlapping ranges)."""
    from alphafive_b200._lib import A5Error
    from alphafive_b200.engine import parse_records
    from alphafive_b200.net import DeviceNet
    S, N, sims = 11, 256, 8
    net = DeviceNet(S, N)
    eng = _engine(S, N, sims, sims + 4, training=True, auto_play=True, seed=2, record_capacity=400)
    prob = torch.empty((N, S * S), device="cuda")
    value = torch.empty((N,), device="cuda")
    eng.step()
    raised = False
    for it in range(4000):
        net.forward_raw(eng.planes_ptr, N, prob, value)
        eng.step(prob, value)
    buf = torch.zeros((400, eng.record_stride), dtype=torch.uint8, device="cuda")
    try:
        eng.harvest(buf)
    except A5Error as err:
        raised = "records lost" in str(err)
    assert raised
    torch.cuda.synchronize()
    recs = parse_records(buf, S)                      # the arena content that was copied before the error
    i = 0
    games = 0
    while i < len(recs) and recs[i]["game_len"] > 0 and i + recs[i]["game_len"] <= 400:
        L = recs[i]["game_len"]
        gid = (recs[i]["game_id"], recs[i]["game_serial"])
        assert [r["ply"] for r in recs[i:i + L]] == list(range(L))
        assert all((r["game_id"], r["game_serial"]) == gid for r in recs[i:i + L])
        i += L
        games += 1
    assert games >= 3 and i > 400 - 121
    assert eng.counters()["records_dropped"] > 0
    eng.close()


@pytest.mark.parametrize("cap", [None, 40])
def test_evaluation_cache_plays_the_same_games(cuda_lib, cap):
    """Cross-game evaluation cache (a5_evalcache_*): leaves served from the table of earlier network results, the
    rest evaluated as a compact batch; with cap < N some leaves wait a pass.  The network is deterministic and
    batch-invariant and every game has its own counter-based RNG stream, so every game must come out bit for bit
    as without the cache -- positions, policies, results -- whatever pass its leaves were evaluated in."""
    from alphafive_b200.engine import parse_records
    from alphafive_b200.net import DeviceNet
    from alphafive_b200.selfplay import SelfPlay
    S, N, sims = 11, 64, 24

    def play(eval_cache, want=70):
        net = DeviceNet(S, N)
        sp = SelfPlay(None, n_games=N, net=net, training=True, seed=4, use_graph=False, eval_cache=eval_cache,
                      board_size=S, simulation_per_step=sims, upper_simulation_per_step=sims + 10)
        games = {}
        while len(games) < want:
            sp.run_passes(400)
            for r in parse_records(sp.harvest()[0], S):
                games.setdefault((r["game_id"], r["game_serial"]), []).append(r)
        stats = sp.cache.stats() if sp.cache is not None else None
        assert sp.engine.counters()["overflows"] == 0
        sp.engine.close(); net.close()
        return games, stats

    plain, _ = play(False)
    cached, stats = play(dict(log2_slots=16, cap=cap) if cap else dict(log2_slots=16), want=100)     # (deferrals shift which games finish first)
    assert stats["hits"] > 0 and stats["stored"] > 0 and (cap is None or stats["deferred"] > 0)
    common = sorted(set(plain) & set(cached))
    assert len(common) >= 50
    for key in common:
        a, b = sorted(plain[key], key=lambda r: r["ply"]), sorted(cached[key], key=lambda r: r["ply"])
        assert len(a) == len(b) == a[0]["game_len"]
        for ra, rb in zip(a, b):
            assert (ra["board"] == rb["board"]).all() and ra["last_action"] == rb["last_action"]
            assert np.array_equal(ra["policy"], rb["policy"]) and ra["value"] == rb["value"] and ra["result"] == rb["result"]


def test_batched_player_with_evaluation_cache_returns_the_same_moves(cuda_lib):
    """BatchedPlayer.get_actions through the evaluation cache (compact batch smaller than the number of players, so
    some leaves wait): policies and actions equal the uncached call's bit for bit, move after move."""
    from alphafive_b200.net import DeviceNet
    from alphafive_b200.selfplay import BatchedPlayer
    S, N, sims = 11, 48, 40
    net = DeviceNet(S, N)
    kw = dict(n_players=N, net=net, training=True, seed=9, board_size=S, simulation_per_step=sims, upper_simulation_per_step=sims + 20)
    a = BatchedPlayer(None, **kw)
    b = BatchedPlayer(None, eval_cache=dict(log2_slots=14, cap=36), **kw)
    boards = np.zeros((N, S, S), np.int8)
    last = np.full(N, -1, np.int32)
    clear = np.ones(N, np.uint8)
    for ply in range(6):
        pa, aa, na, ca = a.get_actions(boards, last, clear=clear, advance=True)
        pb, ab, nb, cb = b.get_actions(boards, last, clear=clear, advance=True)
        assert np.array_equal(aa, ab) and np.array_equal(pa, pb) and np.array_equal(na, nb) and np.array_equal(ca, cb)
        boards, last, clear = na.copy(), aa.copy(), np.zeros(N, np.uint8)
    st = b.cache.stats()
    assert st["hits"] > 0 and st["deferred"] > 0
    a.engine.close(); b.engine.close(); net.close()
