"""Replay-record exchange (the one collective of the path) on CPU: gloo, world_size 2.

Each rank fabricates a different, ragged number of fixed-stride records; every rank must end
up with all of them in rank order, byte-exact, and the parsed games must be the reference's
replay tuples (player.py:77-82 / main.py:94)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alphafive_b200 import replay

S = 11


def _fake_game(gid, serial, L, rng):
    board = np.zeros((S, S), np.int8)
    recs, last = [], -1
    value = 1.0 if L % 2 == 0 else -1.0
    for t in range(L):
        pol = rng.random((S, S)).astype(np.float32)
        pol[board != 0] = 0
        pol /= pol.sum()
        recs.append(dict(game_id=gid, game_serial=serial, ply=t, game_len=L, last_action=last, value=value, weight=0.5 + t,
                         result=1 if L % 2 == 1 else -1, board=board.copy(), policy=pol))
        cell = int(rng.choice(np.flatnonzero(board.reshape(-1) == 0)))
        board.reshape(-1)[cell] = 1
        board = -board
        last, value = cell, -value
    return recs


def _rank_records(rank):
    rng = np.random.default_rng(100 + rank)
    recs = []
    for k in range(2 + 3 * rank):                 # rank 0: 2 games, rank 1: 5 games
        recs += _fake_game(1000 * rank + k, k % 2, 9 + 2 * k + rank, rng)
    return recs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, empty_rank, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        recs = [] if rank == empty_rank else _rank_records(rank)
        local = torch.from_numpy(replay.pack_records(recs, S)) if recs else torch.zeros((0, replay.record_stride(S)), dtype=torch.uint8)
        got, counts = replay.gather_records(local)
        wins = replay.allreduce_wins(rank + 1, 10 * (rank + 1), rank)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), got=got.numpy(), counts=counts.numpy(), wins=np.array(wins))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("empty_rank", [-1, 0])
def test_gather_records_world2(tmp_path, empty_rank):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, empty_rank, str(tmp_path)), nprocs=world, join=True)
    want_parts = [[] if r == empty_rank else _rank_records(r) for r in range(world)]
    want = np.concatenate([replay.pack_records(p, S) if p else np.zeros((0, replay.record_stride(S)), np.uint8) for p in want_parts])
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert z["counts"].tolist() == [len(p) for p in want_parts]
        assert z["got"].shape == want.shape and (z["got"] == want).all()      # same stream on every rank, rank order
        assert z["wins"].tolist() == [3, 30, 1]
    games = replay.records_to_games(replay.parse_records(want, S), S)
    assert len(games) == sum(2 + 3 * r for r in range(world) if r != empty_rank)
    for rec, result in games:
        assert result == (1 if len(rec) % 2 == 1 else -1)
        state, policy, last_action, value, weight = rec[0]
        assert state == "l/" * S and last_action is None and policy.shape == (S, S)
        assert isinstance(value, float) and isinstance(weight, np.float32)


def test_single_process_is_identity():
    local = torch.from_numpy(replay.pack_records(_rank_records(1), S))
    got, counts = replay.gather_records(local)
    assert got is local and counts.tolist() == [local.shape[0]]
    assert replay.record_stride(S) == local.shape[1] and replay.record_stride(15) % 16 == 0
