"""Edge cases of the reference semantics on the device: spent budgets, other win lengths, full boards, tiny boards."""
import numpy as np
import pytest
import torch

from oracle import mcts as omcts, rules as orules

pytestmark = pytest.mark.gpu


def _engine(S, n, sims, upper, **kw):
    from alphafive_b200.engine import SearchEngine, make_config
    return SearchEngine(make_config(board_size=S, simulation_per_step=sims, upper_simulation_per_step=upper, n_games=n, **kw))


def test_budget_rule_when_the_root_is_already_saturated(cuda_lib):
    """player.py:140-143: num = min(sims, upper - sum_n) may be <= 0 -- no simulation is run and calc_policy works on
    the counts the table already holds.  (upper < sims makes the second call on the same root hit that case.)"""
    S, sims, upper = 11, 30, 20
    pv = omcts.table_pv_fn(S, 2)
    rng = np.random.default_rng(0)
    boards = np.stack([orules.random_board(rng, S, 0.15) for _ in range(6)])
    boards = boards[orules.terminal_codes_batch(boards) == 0][:4]
    eng = _engine(S, len(boards), sims, upper, training=False)
    last = np.full(len(boards), -1, np.int32)
    eng.set_roots(boards, last)
    eng.run_search(pv_fn=pv)
    n1 = eng.root_stats()[0].cpu().numpy().copy()
    _, a1 = eng.finish_move()
    eng.set_roots(boards, last)                           # same roots again: sum_n = 29 > upper = 20
    assert int(eng.sims_left().max().item()) <= 0
    eng.run_search(pv_fn=pv)
    n2 = eng.root_stats()[0].cpu().numpy()
    _, a2 = eng.finish_move()
    assert (n1 == n2).all() and torch.equal(a1, a2)
    for j, b in enumerate(boards):
        pl = omcts.OraclePlayer(omcts.SearchConfig(board_size=S, simulation_per_step=sims, upper_simulation_per_step=upper),
                                training=False, pv_fn=pv)
        pl.get_action(b, None)
        _, act = pl.get_action(b, None)
        assert (pl.root_stats(b)[0] == n2[j]).all() and act[0] * S + act[1] == int(a2[j])
    eng.close()


@pytest.mark.parametrize("S,goal", [(7, 4), (9, 6), (11, 3), (15, 5)])
def test_other_win_lengths(cuda_lib, S, goal):
    """config.goal is a parameter of the reference (config.py:6, utils.py:199-235): terminal codes for goal != 5,
    boards with overlines, double wins and full boards included, and a deterministic search that has to see them."""
    from alphafive_b200 import rules
    rng = np.random.default_rng(S * 10 + goal)
    boards = np.stack([orules.random_board(rng, S, f) for f in np.linspace(0.05, 1.0, 400)])
    boards[-1] = np.where(rng.random((S, S)) < 0.5, 1, -1)            # full board
    want = np.array([orules.terminal_code(b, goal) for b in boards])
    got = rules.terminal(torch.from_numpy(boards).cuda(), goal).cpu().numpy()
    assert (got == want).all()
    assert set(want) >= {0, 2} and want[-1] in (1, 2, 3)
    live = boards[want == 0][:6]
    pv = omcts.table_pv_fn(S, 3)
    eng = _engine(S, len(live), 60, 160, goal=goal, training=False)
    eng.set_roots(live, np.full(len(live), -1, np.int32))
    eng.run_search(pv_fn=pv)
    n = eng.root_stats()[0].cpu().numpy()
    for j, b in enumerate(live):
        pl = omcts.OraclePlayer(omcts.SearchConfig(board_size=S, goal=goal, simulation_per_step=60, upper_simulation_per_step=160),
                                training=False, pv_fn=pv)
        pl.get_action(b, None)
        assert (pl.root_stats(b)[0] == n[j]).all(), (S, goal, j)
    eng.close()


def test_tiny_board_games_end_in_draws_and_wins(cuda_lib):
    """5x5 with goal 4: games are short, many fill the board -- exercises the draw label (main.py:88-93: DRAW when the
    last value is 0), restarts and the record arena at the smallest size the ABI accepts."""
    from alphafive_b200.engine import parse_records
    S, N = 5, 64
    eng = _engine(S, N, 12, 20, goal=4, training=True, auto_play=True, seed=5)
    pv = omcts.table_pv_fn(S, 1)
    recs, games = [], 0
    eng.step()
    for it in range(3000):
        need = eng.need_eval().cpu().numpy().astype(bool)
        x = eng.planes().cpu().numpy().astype(np.float32)
        p = np.zeros((N, S * S), np.float32); v = np.zeros((N,), np.float32)
        if need.any():
            p[need], v[need] = pv(x[need])
        eng.step(torch.from_numpy(p).cuda(), torch.from_numpy(v).cuda())
        if it % 100 == 99:
            buf, g_ = eng.harvest()
            recs += parse_records(buf, S); games += g_
            if games >= 60:
                break
    assert games >= 60 and eng.counters()["overflows"] == 0
    by = {}
    for r in recs:
        by.setdefault((r["game_id"], r["game_serial"]), []).append(r)
    results = []
    for plies in by.values():
        plies.sort(key=lambda r: r["ply"])
        L = plies[0]["game_len"]
        assert len(plies) == L and 7 <= L <= 25
        results.append(plies[0]["result"])
        if plies[0]["result"] == 0:
            assert L == 25 and all(r["value"] == 0 for r in plies)
        else:
            assert plies[0]["result"] == (1 if L % 2 == 1 else -1)
    assert 0 in results and (1 in results or -1 in results)
    eng.close()

