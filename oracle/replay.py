"""Oracle (test infrastructure): numpy restatement of the reference's replay buffer,
``utils.RandomStack`` (utils.py:14-146) -- the sink of the finished-game records.

Restated: ``push`` (utils.py:64-116: short-game rejection with probability
``-0.0682 * len + 1.364``, colour re-balancing duplicates, FIFO eviction of the oldest plies
with per-game bookkeeping) and ``get_data`` (utils.py:118-146: sampling without
replacement, the eight board symmetries ``rot90(k)`` + optional ``flip(axis 0)`` with the
``last_action`` remap, ``board_to_inputs``).  The random draws are taken from the two
generators the reference uses -- Python's ``random`` and ``numpy.random`` -- in the
reference's order, so seeding both reproduces its decisions exactly.

Pinned by ``tests/golden/replay_stack.npz``: the real ``RandomStack`` driven by
``oracle.make_golden`` with seeded generators (accept flags, bookkeeping after every push,
one ``get_data`` batch).
"""
from __future__ import annotations

import random as _random

import numpy as np

from . import rules

BLACK_WIN, WHITE_WIN, DRAW = 1, -1, 0


def symmetry_cell(i: int, j: int, k: int, flip: bool, S: int):
    """Where cell (i, j) lands under ``np.rot90(., k, axes=(0, 1))`` then ``np.flip(., 0)``
    -- the reference's ``last_action`` remap (utils.py:132-140)."""
    i, j = [(i, j), (S - 1 - j, i), (S - 1 - i, S - 1 - j), (j, S - 1 - i)][k]
    if flip:
        i = S - 1 - i
    return i, j


class OracleRandomStack:
    """Same state and methods as utils.RandomStack; records are the reference's 5-tuples
    ``(state, policy f32[S,S], last_action | None, value, weight)``."""

    def __init__(self, board_size, length=2000, rnd=None, nprnd=None):
        self.data, self.data_len, self.result = [], [], []
        self.board_size, self.length = board_size, length
        self.white_win = self.black_win = 0
        self.rnd = rnd if rnd is not None else _random          # .random(), .choice()
        self.nprnd = nprnd if nprnd is not None else np.random  # .choice()

    def isEmpty(self):
        return len(self.data) == 0

    def is_full(self):
        return len(self.data) >= self.length

    def push(self, data: list, result: int) -> bool:
        n = len(data)
        if self.rnd.random() <= -0.0682 * n + 1.364:             # utils.py:81
            return False
        self._append(data, result)
        if result == BLACK_WIN:                                  # utils.py:86-92
            self.black_win += 1
            if self.rnd.random() < (self.white_win - self.black_win) / (self.black_win * 1.3):
                self._append(data, result)
                self.black_win += 1
        elif result == WHITE_WIN:                                # utils.py:94-100
            self.white_win += 1
            if self.rnd.random() < (self.black_win - self.white_win) / (self.white_win * 1.02):
                self._append(data, result)
                self.white_win += 1
        beyond = len(self.data) - self.length                    # utils.py:101-115
        if beyond > 0:
            self.data = self.data[beyond:]
            while True:
                if beyond >= self.data_len[0]:
                    beyond -= self.data_len[0]
                    self.data_len.pop(0)
                    r = self.result.pop(0)
                    if r == BLACK_WIN:
                        self.black_win -= 1
                    elif r == WHITE_WIN:
                        self.white_win -= 1
                else:
                    self.data_len[0] -= beyond
                    break
        return True

    def _append(self, data, result):
        self.data.extend(data)
        self.data_len.append(len(data))
        self.result.append(result)

    def draw(self, batch_size):
        """The random choices of one get_data call in the reference's order: indices first, then per
        sample the rotation (numpy stream) and the flip (Python stream)."""
        num = min(batch_size, len(self.data))
        idx = self.nprnd.choice(len(self.data), size=num, replace=False)
        rot = np.empty(num, np.int64)
        flip = np.empty(num, bool)
        for i in range(num):
            rot[i] = self.nprnd.choice([0, 1, 2, 3])
            flip[i] = self.rnd.choice([1, 2]) == 1
        return idx, rot, flip

    def get_data(self, batch_size=1):
        idx, rot, flip = self.draw(batch_size)
        return self.gather(idx, rot, flip)

    def gather(self, idx, rot, flip):
        S = self.board_size
        num = len(idx)
        boards = np.empty((num, 3, S, S), np.float32)
        weights = np.empty((num,), np.float32)
        values = np.empty((num,), np.float32)
        policies = np.empty((num, S, S), np.float32)
        for i, ix in enumerate(idx):
            state, p, la, v, w = self.data[ix]
            board = np.rot90(rules.decode_state(state, S), k=int(rot[i]), axes=(0, 1))
            p = np.rot90(p, k=int(rot[i]), axes=(0, 1))
            if flip[i]:
                board, p = np.flip(board, axis=0), np.flip(p, axis=0)
            if la is not None:
                la = symmetry_cell(la[0], la[1], int(rot[i]), bool(flip[i]), S)
            boards[i] = rules.input_planes(board, la)
            weights[i], values[i], policies[i] = w, v, p
        return boards, weights, values, policies.reshape(num, S * S)
