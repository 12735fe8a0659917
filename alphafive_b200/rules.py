"""Batched rule / encoding kernels (a5_rules_*): torch tensors in, torch tensors out.

Device counterparts of the reference's utils.py:156-283; boards are int8 [n, S, S]
CUDA tensors (+1 side to move, -1 opponent)."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def _boards(boards):
    assert boards.is_cuda and boards.dtype == torch.int8 and boards.dim() == 3
    return boards.contiguous(), boards.shape[0], boards.shape[1]


def terminal(boards, goal=5):
    """utils.is_game_over -> int8 codes: 0 not over, 1 (True,+1), 2 (True,-1), 3 draw."""
    b, n, S = _boards(boards)
    out = torch.empty(n, dtype=torch.int8, device=b.device)
    check(_lib.load().a5_rules_terminal(ptr(b), n, S, goal, ptr(out), stream_ptr()))
    return out


def step(boards, cells):
    """utils.step for flat cells int32[n]."""
    b, n, S = _boards(boards)
    cells = cells.to(torch.int32).contiguous()
    out = torch.empty_like(b)
    check(_lib.load().a5_rules_step(ptr(b), ptr(cells), n, S, ptr(out), stream_ptr()))
    return out


def legal(boards):
    """utils.get_legal_actions -> (mask uint8[n, S*S], count int32[n])."""
    b, n, S = _boards(boards)
    mask = torch.empty((n, S * S), dtype=torch.uint8, device=b.device)
    count = torch.empty(n, dtype=torch.int32, device=b.device)
    check(_lib.load().a5_rules_legal(ptr(b), n, S, ptr(mask), ptr(count), stream_ptr()))
    return mask, count


def inputs(boards, last_cells=None):
    """utils.board_to_inputs -> int8[n, 3, S, S]."""
    b, n, S = _boards(boards)
    if last_cells is not None:
        last_cells = last_cells.to(torch.int32).contiguous()
    out = torch.empty((n, 3, S, S), dtype=torch.int8, device=b.device)
    check(_lib.load().a5_rules_inputs(ptr(b), ptr(last_cells), n, S, ptr(out), stream_ptr()))
    return out


def encode(boards):
    """utils.board_to_state -> list[str]."""
    b, n, S = _boards(boards)
    stride = S * (S + 1) + 4
    buf = torch.zeros((n, stride), dtype=torch.uint8, device=b.device)
    lens = torch.empty(n, dtype=torch.int32, device=b.device)
    check(_lib.load().a5_rules_encode(ptr(b), n, S, ptr(buf), stride, ptr(lens), stream_ptr()))
    host, hl = buf.cpu().numpy(), lens.cpu().numpy()
    return [host[i, :hl[i]].tobytes().decode("ascii") for i in range(n)]


def decode(states, S):
    """utils.state_to_board for a list of state strings -> int8[n, S, S] (CUDA)."""
    import numpy as np
    n = len(states)
    stride = S * (S + 1) + 4
    host = np.zeros((n, stride), np.uint8)
    for i, s in enumerate(states):
        raw = s.encode("ascii")
        assert len(raw) < stride
        host[i, :len(raw)] = np.frombuffer(raw, np.uint8)
    buf = torch.from_numpy(host).cuda()
    out = torch.empty((n, S, S), dtype=torch.int8, device=buf.device)
    check(_lib.load().a5_rules_decode(ptr(buf), stride, n, S, ptr(out), stream_ptr()))
    return out
