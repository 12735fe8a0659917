"""Continuous batching of get_action calls (BatchedPlayer.start_stream / poll / submit over
a5_engine_collect_moves / a5_engine_submit_roots): every player's sequence of (policy, action, next position,
terminal code) must equal, bit for bit, the sequence the lock-step ``get_actions`` call gives it -- only *when* a
search ends differs (player.py:128-147: the players are independent objects)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lockstep_sequences(bp, N, S, moves):
    boards, last, clear = np.zeros((N, S, S), np.int8), np.full(N, -1, np.int32), np.ones(N, np.uint8)
    seq = [[] for _ in range(N)]
    for _ in range(moves):
        pol, act, nxt, codes = bp.get_actions(boards, last, clear=clear, advance=True)
        for g in range(N):
            seq[g].append((pol[g].copy(), int(act[g]), nxt[g].copy(), int(codes[g])))
        over = codes != 0
        boards = np.where(over[:, None, None], 0, nxt).astype(np.int8)      # a finished game restarts (player.py:73)
        last = np.where(over, -1, act).astype(np.int32)
        clear = over.astype(np.uint8)
    return seq


def _stream_sequences(bp, N, S, moves, **kw):
    seq = [[] for _ in range(N)]
    bp.start_stream(np.zeros((N, S, S), np.int8), np.full(N, -1, np.int32), np.ones(N, np.uint8), **kw)
    polls = sizes = 0
    while min(len(s) for s in seq) < moves:
        games, pol, act, nxt, codes = bp.poll()
        polls += 1
        sizes = max(sizes, len(games))
        assert len(set(games.tolist())) == len(games)
        for i, g in enumerate(games):
            seq[g].append((pol[i], int(act[i]), nxt[i], int(codes[i])))
        over = codes != 0
        keep = np.array([len(seq[g]) < moves for g in games], bool)         # players that have all their moves stay parked
        bp.submit(games[keep], np.where(over[:, None, None], 0, nxt).astype(np.int8)[keep],
                  np.where(over, -1, act).astype(np.int32)[keep], over.astype(np.uint8)[keep])
        assert polls < 20000
    return seq, polls, sizes


def _same(a, b, moves):
    for g, (sa, sb) in enumerate(zip(a, b)):
        assert len(sb) == moves, (g, len(sb))
        for m in range(moves):
            assert sa[m][1] == sb[m][1] and sa[m][3] == sb[m][3], (g, m)
            assert np.array_equal(sa[m][0], sb[m][0]) and np.array_equal(sa[m][2], sb[m][2]), (g, m)


@pytest.mark.parametrize("cache", [False, True])
def test_streamed_moves_equal_lockstep_moves(cuda_lib, cache):
    from alphafive_b200.net import DeviceNet
    from alphafive_b200.selfplay import BatchedPlayer
    S, N, sims, moves = 11, 48, 40, 7
    net = DeviceNet(S, N)
    kw = dict(n_players=N, net=net, training=True, seed=9, board_size=S, simulation_per_step=sims,
              upper_simulation_per_step=sims + 20)
    a = BatchedPlayer(None, **kw)
    ref = _lockstep_sequences(a, N, S, moves)
    b = BatchedPlayer(None, eval_cache=dict(log2_slots=14, cap=36) if cache else False, **kw)
    got, polls, sizes = _stream_sequences(b, N, S, moves, passes=4, cap=16)   # cap < N: the first wave of 48 needs 3 polls
    _same(ref, got, moves)
    assert sizes <= 16 and polls > moves
    budgets = {int(r[0].sum() > 0) for s in ref for r in s}
    assert budgets == {1}                                                     # training mode: every call returns a policy
    if cache:
        st = b.cache.stats()
        assert st["hits"] > 0 and st["deferred"] > 0
    a.engine.close(); b.engine.close(); net.close()


def test_streamed_arena_mode_and_parked_players(cuda_lib):
    """training=False, random_a=True (choose_best_player.py:52), 9x9, and players that are never re-submitted stay
    parked: poll returns nothing more for them."""
    from alphafive_b200.net import DeviceNet
    from alphafive_b200.selfplay import BatchedPlayer
    S, N, sims, moves = 9, 32, 24, 4
    net = DeviceNet(S, N)
    kw = dict(n_players=N, net=net, training=False, random_a=True, seed=3, board_size=S, simulation_per_step=sims,
              upper_simulation_per_step=sims + 10)
    a = BatchedPlayer(None, **kw)
    ref = _lockstep_sequences(a, N, S, moves)
    b = BatchedPlayer(None, **kw)
    got, _, _ = _stream_sequences(b, N, S, moves, passes=4)
    _same(ref, got, moves)
    for _ in range(3):                                                        # everybody has `moves` moves and is parked
        games, *_ = b.poll()
        assert len(games) == 0
    assert b.engine.busy() == 0
    a.engine.close(); b.engine.close(); net.close()
