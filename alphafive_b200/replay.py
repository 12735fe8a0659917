"""Replay-record exchange between ranks and conversion to the reference's replay tuples.

Games are independent, so the search path has no collective.  The one exchange
step of the path is the *fill of the shared replay buffer*: every rank harvests the
fixed-stride finished-ply records of its own games (``a5_engine_harvest``) and all ranks
gather all of them, so that each trainer-side ``RandomStack`` sees every game
(main.py:60-61 ``q.get`` -> ``stack.push`` is the single-host analogue).

``gather_records`` works on any ``torch.distributed`` backend: NCCL over NVLink on the GPU
box (device tensors), ``gloo`` in the CPU tests (host tensors).  Counts are exchanged first
(one int64 per rank), then one padded ``all_gather_into_tensor`` moves the payload; the
result keeps rank order, and inside a rank the engine's emission order, so the gathered
stream is deterministic for a given seed and world size.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from ._lib import RecordHeader

HEADER_BYTES = C.sizeof(RecordHeader)


def record_stride(S: int) -> int:
    """Bytes per ply record: header, int8 board padded to 16, f32 policy (a5_record_stride)."""
    Cc = S * S
    body = HEADER_BYTES + (Cc + 15) // 16 * 16 + 4 * Cc
    return (body + 15) // 16 * 16


def gather_records(local: torch.Tensor, group=None, capacity: int | None = None):
    """All ranks' records: ``local`` is uint8 [count, stride] (any count, also 0).

    Returns (records uint8 [total, stride], counts int64 [world]).  ``capacity`` fixes the
    per-rank padded size (records beyond it are a caller error); by default the maximum
    count over ranks is used, which costs one extra tiny all-reduce."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local, torch.tensor([local.shape[0]], dtype=torch.int64)
    world = dist.get_world_size(group)
    dev = local.device
    stride = local.shape[1]
    cnt = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    counts = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, cnt, group=group)
    if capacity is None:
        capacity = int(counts.max().item())
    assert int(counts.max().item()) <= capacity, "record capacity too small for this harvest"
    if capacity == 0:
        return local[:0], counts.cpu()
    send = torch.zeros((capacity, stride), dtype=torch.uint8, device=dev)
    send[:local.shape[0]] = local
    recv = torch.empty((world, capacity, stride), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
    host_counts = counts.cpu()
    parts = [recv[r, :int(host_counts[r])] for r in range(world)]
    return torch.cat(parts, dim=0), host_counts


def allreduce_wins(wins0: int, wins1: int, draws: int, group=None, device=None):
    """Arena bookkeeping across ranks: one sum all-reduce of three counters."""
    t = torch.tensor([wins0, wins1, draws], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tuple(int(x) for x in t.cpu())


# ---------------------------------------------------------------------------------------
# record bytes -> the reference's replay tuples (player.py:77-82) and back
# ---------------------------------------------------------------------------------------
HEADER_DTYPE = np.dtype([("game_id", "<i8"), ("game_serial", "<i4"), ("ply", "<i2"), ("game_len", "<i2"),
                         ("last_action", "<i4"), ("value", "<f4"), ("weight", "<f4"), ("result", "<i4")])
assert HEADER_DTYPE.itemsize == HEADER_BYTES


def parse_headers(buf) -> np.ndarray:
    """The a5_record_header of every record as one structured array (one D2H copy of 32 B per record)."""
    if isinstance(buf, torch.Tensor):
        buf = buf[:, :HEADER_BYTES].contiguous().cpu().numpy()
    else:
        buf = np.ascontiguousarray(np.asarray(buf)[:, :HEADER_BYTES])
    return buf.view(HEADER_DTYPE).reshape(-1)


def parse_records(buf, S: int):
    """uint8 [count, stride] (torch or numpy) -> list of dicts with numpy fields."""
    host = buf.cpu().numpy() if isinstance(buf, torch.Tensor) else np.asarray(buf)
    Cc = S * S
    bb = (Cc + 15) // 16 * 16
    n = host.shape[0]
    if n == 0:
        return []
    head = parse_headers(host)
    boards = np.ascontiguousarray(host[:, HEADER_BYTES:HEADER_BYTES + Cc]).view(np.int8).reshape(n, S, S)
    policies = np.ascontiguousarray(host[:, HEADER_BYTES + bb:HEADER_BYTES + bb + 4 * Cc]).view(np.float32).reshape(n, S, S)
    cols = {k: head[k].tolist() for k in HEADER_DTYPE.names}
    return [dict(game_id=cols["game_id"][i], game_serial=cols["game_serial"][i], ply=cols["ply"][i],
                 game_len=cols["game_len"][i], last_action=cols["last_action"][i],
                 value=np.float32(cols["value"][i]).item(), weight=np.float32(cols["weight"][i]).item(),
                 result=cols["result"][i], board=boards[i], policy=policies[i]) for i in range(n)]


def pack_records(recs, S: int) -> np.ndarray:
    """Inverse of ``parse_records`` (tests and tools)."""
    Cc = S * S
    bb = (Cc + 15) // 16 * 16
    stride = record_stride(S)
    out = np.zeros((len(recs), stride), np.uint8)
    for i, r in enumerate(recs):
        h = RecordHeader(game_id=r["game_id"], game_serial=r["game_serial"], ply=r["ply"], game_len=r["game_len"],
                         last_action=r["last_action"], value=r["value"], weight=r["weight"], result=r["result"])
        out[i, :HEADER_BYTES] = np.frombuffer(bytes(h), np.uint8)
        out[i, HEADER_BYTES:HEADER_BYTES + Cc] = np.asarray(r["board"], np.int8).reshape(-1).view(np.uint8)
        out[i, HEADER_BYTES + bb:HEADER_BYTES + bb + 4 * Cc] = np.asarray(r["policy"], np.float32).reshape(-1).view(np.uint8)
    return out


def records_to_games(recs, S: int):
    """Finished games as ``(record list, result)`` like ``q.put((game_record, result))``
    (main.py:94), each record ``(state, policy, last_action, value, weight)``."""
    from .genData.player import board_to_state
    by = {}
    for r in recs:
        by.setdefault((r["game_id"], r["game_serial"]), []).append(r)
    games = []
    for key in sorted(by):
        plies = sorted(by[key], key=lambda r: r["ply"])
        rec = []
        for r in plies:
            la = None if r["last_action"] < 0 else (r["last_action"] // S, r["last_action"] % S)
            rec.append((board_to_state(r["board"]), r["policy"], la, float(r["value"]), np.float32(r["weight"])))
        games.append((rec, int(plies[0]["result"])))
    return games
