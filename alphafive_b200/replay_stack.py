"""Device-resident replay buffer with the surface of the reference's ``utils.RandomStack``
(utils.py:14-146): ``push / get_data / is_full / isEmpty / save / load`` and the counters the
trainer prints (``main.py:34,60-78``).

The plies live in HBM as a ring of the engine's fixed-stride records (``a5_record_header`` +
int8 board + f32 policy) -- what ``a5_engine_harvest`` and the NCCL gather of
``alphafive_b200.replay`` deliver -- so a finished game goes from the search to a training batch
without passing through host memory.  Host side only the per-game bookkeeping of the reference
stays (``data_len``, ``result``, colour counters: a few integers per game), together with its
random decisions, drawn from the same two generators in the same order (Python ``random`` for
rejection / duplication / flips, ``numpy.random`` for the sample indices and rotations), so that
seeding both reproduces the reference's behaviour exactly.  ``get_data`` runs the gather, the
eight-fold symmetry augmentation and ``board_to_inputs`` in one CUDA kernel
(``a5_replay_sample``, csrc/replay.cu).
"""
from __future__ import annotations

import os
import pickle
import random
from time import time

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .replay import HEADER_BYTES, pack_records, parse_headers, parse_records, record_stride

BLACK_WIN, WHITE_WIN, DRAW = 1, -1, 0


class RandomStack:
    def __init__(self, board_size, length=2000, device=None):
        self.lib = _lib.load()
        self.board_size = board_size
        self.length = length
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.stride = record_stride(board_size)
        # a push may hold up to two copies of one game beyond `length` before the oldest plies are evicted
        self.capacity = length + 2 * board_size * board_size
        self.ring = torch.zeros((self.capacity, self.stride), dtype=torch.uint8, device=self.device)
        self.head = 0                  # ring row of the oldest ply
        self.count = 0                 # plies held == len(self.data) of the reference
        self.white_win = self.black_win = 0
        self.data_len, self.result = [], []
        self.total_length = self.num = 0
        self.time = time()
        self.self_play_black_win = self.self_play_white_win = 0

    # -- reference surface -----------------------------------------------------------------
    def __len__(self):
        return self.count

    def isEmpty(self):
        return self.count == 0

    def is_full(self):
        return self.count >= self.length

    def push(self, data, result: int) -> bool:
        """utils.py:64-116.  ``data``: a game as the reference's list of 5-tuples
        ``(state, policy, last_action, value, weight)`` or as uint8 [plies, stride] records (device or
        host tensor / array)."""
        rec = self._as_records(data, result)
        n = rec.shape[0]
        self.total_length += n
        self.num += 1
        if result == BLACK_WIN:
            self.self_play_black_win += 1
        elif result == WHITE_WIN:
            self.self_play_white_win += 1
        if self.total_length >= 100:                            # the reference prints these (utils.py:72-79)
            self.total_length = self.num = 0
            self.time = time()
        if random.random() <= -0.0682 * n + 1.364:             # utils.py:81: short games are dropped
            return False
        self._append(rec, result)
        if result == BLACK_WIN:                                 # utils.py:86-100: colour re-balancing
            self.black_win += 1
            if random.random() < (self.white_win - self.black_win) / (self.black_win * 1.3):
                self._append(rec, result)
                self.black_win += 1
        elif result == WHITE_WIN:
            self.white_win += 1
            if random.random() < (self.black_win - self.white_win) / (self.white_win * 1.02):
                self._append(rec, result)
                self.white_win += 1
        beyond = self.count - self.length                       # utils.py:101-115: FIFO eviction
        if beyond > 0:
            self.head = (self.head + beyond) % self.capacity
            self.count -= beyond
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

    def push_records(self, records, S=None):
        """All finished games of a harvest / gather (uint8 [count, stride], any order of games; plies of
        a game are contiguous and ordered as the engine emits them).  Returns the accept flags."""
        if records.shape[0] == 0:
            return []
        head = parse_headers(records)                           # one 32 B/record D2H copy for the whole harvest
        starts = np.flatnonzero(head["ply"] == 0)
        lens, res = head["game_len"][starts].tolist(), head["result"][starts].tolist()
        assert sum(lens) == records.shape[0], "records are not whole games with contiguous plies"
        return [self.push(records[i:i + n], r) for i, n, r in zip(starts.tolist(), lens, res)]

    def get_data(self, batch_size=1):
        """utils.py:118-146 -> (boards f32[n,3,S,S], weights f32[n], values f32[n], policies f32[n,S*S])
        as numpy arrays, like the reference."""
        b, w, v, p = self.get_data_device(batch_size)
        return b.cpu().numpy(), w.cpu().numpy(), v.cpu().numpy(), p.cpu().numpy()

    def get_data_device(self, batch_size=1):
        """Same batch as device tensors (for a trainer that stays on the GPU)."""
        S = self.board_size
        num = min(batch_size, self.count)
        idx = np.random.choice(self.count, size=num, replace=False)
        # utils.py:128,136 draw a rotation (numpy stream) and a flip (Python stream) per sample.  The two
        # streams are independent, and legacy np.random.choice over a 4-element list is one bounded 32-bit draw,
        # so one bulk randint consumes the numpy stream exactly like the reference's per-sample calls
        # (tests/test_gpu_replay_stack.py compares batches with the real RandomStack under equal seeds)
        rot = np.random.randint(0, 4, size=num).astype(np.uint8)
        flip = np.fromiter((random.choice((1, 2)) == 1 for _ in range(num)), np.uint8, num)
        return self.gather((self.head + idx.astype(np.int64)) % self.capacity, rot, flip)

    def gather(self, rows, rot, flip):
        S, dev = self.board_size, self.device
        num = len(rows)
        d_idx = torch.from_numpy(np.ascontiguousarray(rows, np.int64)).to(dev)
        d_rot = torch.from_numpy(np.ascontiguousarray(rot, np.uint8)).to(dev)
        d_flip = torch.from_numpy(np.ascontiguousarray(flip, np.uint8)).to(dev)
        boards = torch.empty((num, 3, S, S), dtype=torch.float32, device=dev)
        weights = torch.empty((num,), dtype=torch.float32, device=dev)
        values = torch.empty((num,), dtype=torch.float32, device=dev)
        policies = torch.empty((num, S * S), dtype=torch.float32, device=dev)
        check(self.lib.a5_replay_sample(ptr(self.ring), S, ptr(d_idx), ptr(d_rot), ptr(d_flip), num, ptr(boards),
                                        ptr(weights), ptr(values), ptr(policies), stream_ptr()))
        return boards, weights, values, policies

    # -- the reference's pickle files (utils.py:29-57), for interchange with its trainer --------
    @property
    def data(self):
        """The held plies as the reference's list of 5-tuples (copies the ring to the host)."""
        from .genData.player import board_to_state
        S = self.board_size
        out = []
        for r in parse_records(self._linear(), S):
            la = None if r["last_action"] < 0 else (r["last_action"] // S, r["last_action"] % S)
            out.append((board_to_state(r["board"]), r["policy"], la, float(r["value"]), np.float32(r["weight"])))
        return out

    def save(self, s="", directory="data_buffer"):
        os.makedirs(directory, exist_ok=True)
        for name, obj in (("data", self.data), ("data_len", list(self.data_len)), ("result", list(self.result))):
            with open(os.path.join(directory, f"{name}{s}.pkl"), "wb") as f:
                pickle.dump(obj, f)

    def load(self, s="", directory="data_buffer"):
        with open(os.path.join(directory, f"data{s}.pkl"), "rb") as f:
            data = pickle.load(f)
        with open(os.path.join(directory, f"data_len{s}.pkl"), "rb") as f:
            data_len = pickle.load(f)
        with open(os.path.join(directory, f"result{s}.pkl"), "rb") as f:
            result = pickle.load(f)
        if len(data) > self.capacity:
            raise ValueError(f"{len(data)} plies do not fit a RandomStack of length {self.length}")
        rec = self._tuples_to_records(data, 0)
        self.ring[:rec.shape[0]] = rec.to(self.device)
        self.head, self.count = 0, rec.shape[0]
        self.data_len, self.result = list(data_len), list(result)
        self.white_win = self.result.count(WHITE_WIN)
        self.black_win = self.result.count(BLACK_WIN)

    # -- internals -------------------------------------------------------------------------
    def _tuples_to_records(self, data, result):
        from .genData.player import state_to_board
        S = self.board_size
        recs = []
        for ply, (state, policy, la, value, weight) in enumerate(data):
            recs.append(dict(game_id=0, game_serial=0, ply=ply & 0x7fff, game_len=min(len(data), 0x7fff),
                             last_action=-1 if la is None else int(la[0]) * S + int(la[1]), value=float(value),
                             weight=float(weight), result=int(result), board=state_to_board(state, S),
                             policy=np.asarray(policy, np.float32)))
        return torch.from_numpy(pack_records(recs, S))

    def _as_records(self, data, result):
        if isinstance(data, torch.Tensor):
            rec = data
        elif isinstance(data, np.ndarray):
            rec = torch.from_numpy(data)
        else:
            rec = self._tuples_to_records(list(data), result)
        assert rec.dtype == torch.uint8 and rec.dim() == 2 and rec.shape[1] == self.stride, "bad record array"
        return rec

    def _append(self, rec, result):
        n = rec.shape[0]
        assert self.count + n <= self.capacity, "game longer than the ring's slack"
        tail = (self.head + self.count) % self.capacity
        first = min(n, self.capacity - tail)
        self.ring[tail:tail + first].copy_(rec[:first], non_blocking=True)
        if first < n:
            self.ring[:n - first].copy_(rec[first:], non_blocking=True)
        self.count += n
        self.data_len.append(n)
        self.result.append(result)

    def _linear(self):
        rows = (self.head + torch.arange(self.count, device=self.device)) % self.capacity
        return self.ring[rows]
