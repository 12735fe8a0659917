"""Two-half pipeline on SM-partitioned streams (alphafive_b200.pipeline) against the single-stream
schedule: same seeds and global game ids -> bit-identical finished-game records and counters."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _records(buf):
    a = buf.cpu().numpy()
    key = a[:, :16].copy().view(np.int64)                 # (game_id, serial|ply|len) sorts games and plies
    order = np.lexsort((key[:, 1], key[:, 0]))
    return a[order]


def test_pipelined_selfplay_equals_single_stream(cuda_lib):
    from alphafive_b200.net import DeviceNet, glorot_init
    from alphafive_b200.selfplay import PipelinedSelfPlay, SelfPlay
    S, N, passes = 11, 256, 700
    w = glorot_init(S, 3)
    kw = dict(board_size=S, simulation_per_step=12, upper_simulation_per_step=16)
    one = SelfPlay(None, n_games=N, net=DeviceNet(S, N, w), training=True, seed=5, game_id_base=1000, use_graph=False, **kw)
    one.run_passes(passes)
    want, want_games = one.harvest()
    want = _records(want)
    two = PipelinedSelfPlay(None, n_games=N, weights=w, training=True, seed=5, game_id_base=1000, **kw)
    assert two.pipe.part.n_small + two.pipe.part.n_big <= torch.cuda.get_device_properties(0).multi_processor_count
    two.run_passes(passes)
    got, got_games = two.harvest()
    got = _records(got)
    assert want_games == got_games and want_games > 20
    assert want.shape == got.shape and (want == got).all()
    c1, c2 = one.counters(), two.counters()
    for k in ("moves", "sims", "leaf_evals", "games", "selects", "overflows"):
        assert c1[k] == c2[k], k
    # and it keeps going after a harvest / budget change
    two.set_budget(8, 10)
    two.run_passes(50)
    assert two.counters()["moves"] > c2["moves"]


def test_pipelined_batched_player_equals_batched_player(cuda_lib):
    """Host-buffer API: same policies and actions from the pipelined and the single-stream search (two moves,
    tree reuse in between)."""
    from alphafive_b200.net import DeviceNet, glorot_init
    from alphafive_b200.selfplay import BatchedPlayer, PipelinedBatchedPlayer
    from oracle import rules as orules
    S, N = 11, 64
    w = glorot_init(S, 1)
    rng = np.random.default_rng(2)
    boards = np.stack([orules.random_board(rng, S, 0.05 * (i % 5)) for i in range(N)])
    for i in range(N):
        while orules.terminal(boards[i])[0]:
            boards[i] = orules.random_board(rng, S, 0.1)
    last = np.full(N, -1, np.int32)
    kw = dict(board_size=S, simulation_per_step=40, upper_simulation_per_step=50, max_inner=16)
    a = BatchedPlayer(None, n_players=N, net=DeviceNet(S, N, w), training=True, seed=9, game_id_base=50, **kw)
    b = PipelinedBatchedPlayer(None, n_players=N, weights=w, training=True, seed=9, game_id_base=50, **kw)
    clear = np.ones(N, np.uint8)
    for move in range(2):
        pa, aa, na, ca = a.get_actions(boards, last, None, clear, advance=True)
        pb, ab, nb, cb = b.get_actions(boards, last, None, clear, advance=True)
        assert (aa == ab).all() and (pa == pb).all() and (na == nb).all() and (ca == cb).all(), move
        boards, last, clear = na.copy(), aa.copy(), np.zeros(N, np.uint8)
