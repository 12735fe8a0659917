"""Board sizes other than the two BASELINE.json names (config.board_size is a free parameter of the reference,
config.py:2; A5_MAX_BOARD = 15): rules, the three net paths and the deterministic search against the oracle.
The oracle itself is pinned to the reference at 11 and 15 (tests/golden); these sizes check that nothing on the
device depends on (S + 1)^2 being a multiple of the tile sizes."""
import numpy as np
import pytest
import torch

from oracle import mcts as omcts, net as onet, rules as orules

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("S", [7, 9, 13])
def test_other_board_sizes(cuda_lib, S):
    from alphafive_b200 import _lib, rules
    from alphafive_b200.engine import SearchEngine, make_config
    from alphafive_b200.net import DeviceNet, glorot_init
    rng = np.random.default_rng(S)
    boards = np.stack([orules.random_board(rng, S, 0.05 * (i % 9)) for i in range(257)])
    # rules: terminal codes, legal masks and codecs bit-exact
    dev = torch.from_numpy(boards).cuda()
    assert (rules.terminal(dev).cpu().numpy() == orules.terminal_codes_batch(boards)).all()
    live = boards[orules.terminal_codes_batch(boards) == 0][:70]
    # net: every compute path within 1e-4 of the fp32 oracle (ragged batch: 70 boards through max_batch 32)
    w = glorot_init(S, seed=1)
    x = np.stack([orules.input_planes(b) for b in live])
    want_p, want_v = onet.OracleNet(S, w).eval(x)
    net = DeviceNet(S, 32, w)
    for mode in (_lib.NET_TC, _lib.NET_FP32):
        net.mode = mode
        p, v = net.eval(x)
        assert np.abs(p - want_p).max() <= 1e-4 and np.abs(v - want_v).max() <= 1e-4, (S, mode)
    xs = torch.from_numpy(np.ascontiguousarray(x > 0.5)).to(torch.int8).cuda()
    for i, n in ((0, 1), (1, 3), (4, 8)):
        p, v = net.forward(xs[i:i + n], mode=_lib.NET_SMALL)
        assert np.abs(p.cpu().numpy() - want_p[i:i + n]).max() <= 2e-5, (S, n)
        assert np.abs(v.cpu().numpy() - want_v[i:i + n]).max() <= 2e-5, (S, n)
    # deterministic search under a tie-free table pv_fn: visit counts equal the oracle's
    sims, N = 40, 12
    pv = omcts.table_pv_fn(S, 5)
    eng = SearchEngine(make_config(board_size=S, simulation_per_step=sims, upper_simulation_per_step=sims + 100,
                                   n_games=N, training=False))
    eng.set_roots(live[:N], np.full(N, -1, np.int32))
    eng.run_search(pv_fn=pv)
    n_dev = eng.root_stats()[0].cpu().numpy()
    for j in range(N):
        pl = omcts.OraclePlayer(omcts.SearchConfig(board_size=S, simulation_per_step=sims, upper_simulation_per_step=sims + 100),
                                training=False, pv_fn=pv)
        pl.get_action(live[j], None)
        assert (pl.root_stats(live[j])[0] == n_dev[j]).all(), (S, j)
    assert eng.counters()["overflows"] == 0
    eng.close()
