"""CUDA rule kernels (through the C ABI) vs the reference's golden vectors and the oracle."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import rules as orules

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("S", [11, 15])
def test_rules_golden(cuda_lib, S):
    from alphafive_b200 import rules
    g = golden(f"rules_{S}.npz")
    boards = torch.from_numpy(g["boards"]).cuda()
    n = boards.shape[0]
    assert (rules.terminal(boards).cpu().numpy() == g["codes"]).all()
    assert rules.encode(boards) == [str(s) for s in g["states"]]
    assert (rules.decode([str(s) for s in g["states"]], S).cpu().numpy() == g["boards"]).all()
    mask, count = rules.legal(boards)
    assert (count.cpu().numpy() == g["nlegal"]).all()
    assert (mask.cpu().numpy().reshape(n, S, S) == (g["boards"] == 0)).all()
    la = g["last_action"].astype(np.int64)
    last = np.where(la[:, 0] >= 0, la[:, 0] * S + la[:, 1], -1).astype(np.int32)
    planes = rules.inputs(boards, torch.from_numpy(last).cuda())
    assert (planes.cpu().numpy() == g["inputs"]).all()
    act = g["actions"].astype(np.int64)
    ok = act[:, 0] >= 0
    cells = torch.from_numpy((act[ok, 0] * S + act[ok, 1]).astype(np.int32)).cuda()
    stepped = rules.step(boards[torch.from_numpy(ok).cuda()], cells)
    assert (stepped.cpu().numpy() == g["stepped"][ok]).all()


@pytest.mark.parametrize("S,n", [(11, 1_000_000), (15, 300_000)])
def test_terminal_million_random_boards(cuda_lib, S, n):
    from alphafive_b200 import rules
    rng = np.random.default_rng(S)
    fill = rng.uniform(0, 1, size=(n, 1, 1))
    u = rng.random((n, S, S))
    boards = np.zeros((n, S, S), np.int8)
    boards[u < fill / 2] = 1
    boards[(u >= fill / 2) & (u < fill)] = -1
    got = rules.terminal(torch.from_numpy(boards).cuda()).cpu().numpy()
    want = np.concatenate([orules.terminal_codes_batch(boards[i:i + 100_000]) for i in range(0, n, 100_000)])
    assert (got == want).all()
    assert len(np.unique(want)) >= 3


def test_encode_decode_round_trip_and_edge_cases(cuda_lib):
    from alphafive_b200 import rules
    S = 11
    empty = np.zeros((1, S, S), np.int8)
    full = np.fromfunction(lambda _, i, j: ((i // 2 + j) % 2) * 2 - 1, (1, S, S)).astype(np.int8)
    boards = torch.from_numpy(np.concatenate([empty, full])).cuda()
    st = rules.encode(boards)
    assert st[0] == "l/" * 11 and len(st[1]) == S * (S + 1)
    assert (rules.decode(st, S).cpu().numpy() == boards.cpu().numpy()).all()
    codes = rules.terminal(boards).cpu().numpy()
    assert codes[0] == 0 and codes[1] == orules.terminal_code(full[0])
    assert rules.terminal(boards[:0]).numel() == 0            # empty batch
