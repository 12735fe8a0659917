"""Free functions with the signatures of the reference's ``utils.py:149-296`` -- what the
drivers import (``self_play.py:81,98-101``, ``choose_best_player.py:53-56``, ``main.py:88-93``):
``state_to_board, board_to_state, step, is_game_over, get_legal_actions, board_to_inputs,
construct_weights`` and the result constants.

Single boards in, single results out, exactly as the reference; the rule work itself runs in
the batched CUDA kernels (``a5_rules_*``, one warp per board) with a batch of one.  For many
boards at once use ``alphafive_b200.rules`` directly.  The string <-> board conversions are
host-side parsing at the API boundary (strings never reach the device search).
"""
from __future__ import annotations

import numpy as np
import torch

from . import rules as _rules
from .genData.player import board_to_state, construct_weights, state_to_board  # noqa: F401  (re-exported)
from .replay_stack import RandomStack  # noqa: F401  (utils.py:14-146, device-resident)

BLACK_WIN = 1          # utils.py:9-11
WHITE_WIN = -1
DRAW = 0

_CODE = {0: (False, 0.0), 1: (True, 1.0), 2: (True, -1.0), 3: (True, 0.0)}


def _dev(board) -> torch.Tensor:
    b = np.ascontiguousarray(np.asarray(board, dtype=np.int8))
    assert b.ndim == 2 and b.shape[0] == b.shape[1], "board must be int8[S, S]"
    return torch.from_numpy(b[None]).cuda()


def is_game_over(board: np.ndarray, goal: int) -> tuple:
    """utils.py:199-235 -> (over: bool, value: float) in the reference's scan order."""
    return _CODE[int(_rules.terminal(_dev(board), goal).cpu()[0])]


def get_legal_actions(board: np.ndarray):
    """utils.py:238-245 -> [(i, j), ...] row-major."""
    S = np.asarray(board).shape[0]
    mask, _ = _rules.legal(_dev(board))
    cells = np.flatnonzero(mask.cpu().numpy()[0])
    return [(int(c) // S, int(c) % S) for c in cells]


def board_to_inputs(board: np.ndarray, type_=np.float32, last_action=None):
    """utils.py:256-272 -> type_[3, S, S] planes (own, opponent, one-hot last move)."""
    S = np.asarray(board).shape[0]
    last = None
    if last_action is not None:
        last = torch.tensor([int(last_action[0]) * S + int(last_action[1])], dtype=torch.int32, device="cuda")
    return _rules.inputs(_dev(board), last).cpu().numpy()[0].astype(type_)


def step(board: np.ndarray, action: tuple):
    """utils.py:275-283: places +1 at ``action`` IN PLACE (as the reference does) and returns
    the negated board as a new array."""
    S = board.shape[0]
    cell = torch.tensor([int(action[0]) * S + int(action[1])], dtype=torch.int32, device="cuda")
    out = _rules.step(_dev(board), cell).cpu().numpy()[0]
    board[action[0], action[1]] = 1
    return out.astype(board.dtype, copy=False)


def softmax(x):
    """utils.py:149-153 (host helper, not on the hot path)."""
    probs = np.exp(x - np.max(x))
    return probs / np.sum(probs)
