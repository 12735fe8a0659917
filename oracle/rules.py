"""Oracle (test infrastructure): Gomoku rules and encodings, numpy restatement.

Restates the behaviour of the reference's ``utils.py`` rule functions.  Boards
are ``int8[S, S]`` with +1 = side to move, -1 = opponent, 0 = empty; after every
move the board is negated so the side to move is always +1 (utils.py:275-283).

Pinned against the reference itself by ``tests/golden/rules_*.npz`` (written by
``oracle/make_golden.py`` from the imported reference functions).
"""
from __future__ import annotations

import numpy as np

# outcome codes used by the trainer side (utils.py:9-11)
BLACK_WIN, WHITE_WIN, DRAW = 1, -1, 0


# --------------------------------------------------------------------------- #
# state string  <->  board   (utils.py:156-175 / 178-196)
# --------------------------------------------------------------------------- #
def encode_state(board: np.ndarray) -> str:
    """Row-wise run-length string: a run of k>=1 empties is chr('a'+k), a stone is
    str(v+2) ('3' = side to move, '1' = opponent), each row ends with '/'.
    Restates utils.py:156-175 (board_to_state)."""
    out = []
    for row in np.asarray(board):
        run = 0
        for v in row.tolist():
            if v == 0:
                run += 1
                continue
            if run:
                out.append(chr(97 + run))
                run = 0
            out.append("3" if v > 0 else "1")
        if run:
            out.append(chr(97 + run))
        out.append("/")
    return "".join(out)


def decode_state(state: str, size: int) -> np.ndarray:
    """Inverse of :func:`encode_state` (utils.py:178-196, state_to_board)."""
    board = np.zeros((size, size), np.int8)
    r = c = 0
    for ch in state:
        if ch == "/":
            r, c = r + 1, 0
        elif "a" <= ch <= "z" or "A" <= ch <= "Z":
            c += ord(ch) - 97
        else:
            board[r, c] = int(ch) - 2
            c += 1
    return board


def initial_state(size: int) -> str:
    """Empty board string, e.g. 'l/'*11 (player.py:37-46)."""
    return (chr(97 + size) + "/") * size


# --------------------------------------------------------------------------- #
# terminal test   (utils.py:199-235)
# --------------------------------------------------------------------------- #
def window_sums(board: np.ndarray, goal: int = 5) -> np.ndarray:
    """``int[S, S, 4]`` sums of the ``goal``-long windows anchored at (i, j) in the
    reference's four directions, in its per-cell order: 0 = down ``(i+k, j)``,
    1 = right ``(i, j+k)``, 2 = down-right ``(i+k, j+k)``, 3 = up-right ``(i-k, j+k)``.
    Windows clipped by the edge are shorter in the reference (utils.py:210,215) and
    can never reach +-goal, and the diagonals are simply skipped there
    (utils.py:221,227); both are represented here by 0."""
    b = np.asarray(board).astype(np.int32)
    S = b.shape[0]
    out = np.zeros((S, S, 4), np.int32)
    n = S - goal + 1
    if n <= 0:
        return out
    k = np.arange(goal)
    out[:n, :, 0] = sum(b[t:t + n, :] for t in k)
    out[:, :n, 1] = sum(b[:, t:t + n] for t in k)
    out[:n, :n, 2] = sum(b[t:t + n, t:t + n] for t in k)
    out[goal - 1:, :n, 3] = sum(b[goal - 1 - t:goal - 1 - t + n, t:t + n] for t in k)
    return out


def terminal(board: np.ndarray, goal: int = 5):
    """(over, value) from the point of view of the side to move, *before* it moves.

    The reference scans cells row-major and, per cell, the four directions in the
    order of :func:`window_sums`; the first window that sums to +goal returns
    (True, 1.0), to -goal (True, -1.0) (utils.py:207-232).  The scan order only
    matters for boards holding fives of both colours.  Then a full board is a draw
    (True, 0.0) (utils.py:233-234); otherwise (False, 0.0).  Overlines count."""
    ws = window_sums(board, goal).reshape(-1)
    hit = np.flatnonzero(np.abs(ws) == goal)
    if hit.size:
        return True, (1.0 if ws[hit[0]] > 0 else -1.0)
    if not (np.asarray(board) == 0).any():
        return True, 0.0
    return False, 0.0


def terminal_code(board: np.ndarray, goal: int = 5) -> int:
    """Compact code used by the device kernels: 0 = not over, 1 = (True, +1.0),
    2 = (True, -1.0), 3 = draw (True, 0.0)."""
    over, v = terminal(board, goal)
    if not over:
        return 0
    return 1 if v > 0 else 2 if v < 0 else 3


# --------------------------------------------------------------------------- #
# moves, inputs, training weights
# --------------------------------------------------------------------------- #
def legal_cells(board: np.ndarray) -> np.ndarray:
    """Flat indices ``i*S+j`` of the empty cells in row-major order -- the order of
    utils.py:238-245 (get_legal_actions), which fixes edge order in every node."""
    return np.flatnonzero(np.asarray(board).reshape(-1) == 0)


def legal_actions(board: np.ndarray):
    S = board.shape[1]
    return [(int(c) // S, int(c) % S) for c in legal_cells(board)]


def play(board: np.ndarray, action) -> np.ndarray:
    """Place a +1 stone at ``action`` and hand the board to the opponent (negate).
    Unlike utils.py:275-283 (step) the input is not mutated."""
    nxt = np.array(board, dtype=np.int8, copy=True)
    nxt[action[0], action[1]] = 1
    return (-nxt).astype(np.int8)


def input_planes(board: np.ndarray, last_action=None, dtype=np.float32) -> np.ndarray:
    """``[3, S, S]`` network input: own stones, opponent stones, one-hot of the
    previous move (all zero when ``last_action`` is None).  utils.py:256-272."""
    b = np.asarray(board)
    x = np.zeros((3,) + b.shape, dtype)
    x[0][b == 1] = 1
    x[1][b == -1] = 1
    if last_action is not None:
        x[2, last_action[0], last_action[1]] = 1
    return x


def ply_weights(length: int, gamma: float = 0.95) -> np.ndarray:
    """Per-ply training weights: geometric in float32 from the last ply backwards,
    rescaled to sum to ``length`` (utils.py:286-296, construct_weights)."""
    w = np.empty((int(length),), np.float32)
    w[-1] = 1.0
    for i in range(length - 2, -1, -1):
        w[i] = w[i + 1] * gamma
    return length * w / np.sum(w)


def random_board(rng: np.random.Generator, size: int, fill: float | None = None) -> np.ndarray:
    """Random test board with roughly ``fill`` of the cells occupied."""
    if fill is None:
        fill = rng.uniform(0.0, 1.0)
    u = rng.random((size, size))
    b = np.zeros((size, size), np.int8)
    b[u < fill / 2] = 1
    b[(u >= fill / 2) & (u < fill)] = -1
    return b


def terminal_codes_batch(boards: np.ndarray, goal: int = 5) -> np.ndarray:
    """Vectorised :func:`terminal_code` over ``int8[N, S, S]`` (same scan-order rule)."""
    b = np.asarray(boards).astype(np.int32)
    N, S, _ = b.shape
    n = S - goal + 1
    ws = np.zeros((N, S, S, 4), np.int32)
    k = range(goal)
    ws[:, :n, :, 0] = sum(b[:, t:t + n, :] for t in k)
    ws[:, :, :n, 1] = sum(b[:, :, t:t + n] for t in k)
    ws[:, :n, :n, 2] = sum(b[:, t:t + n, t:t + n] for t in k)
    ws[:, goal - 1:, :n, 3] = sum(b[:, goal - 1 - t:goal - 1 - t + n, t:t + n] for t in k)
    flat = ws.reshape(N, -1)
    hit = np.abs(flat) == goal
    first = hit.argmax(1)
    anyhit = hit.any(1)
    sign = flat[np.arange(N), first]
    full = ~(b == 0).reshape(N, -1).any(1)
    return np.where(anyhit, np.where(sign > 0, 1, 2), np.where(full, 3, 0)).astype(np.int8)
