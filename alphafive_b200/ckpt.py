"""TensorFlow checkpoint-v2 ("bundle") reader, without TensorFlow.

Replaces the loading half of the reference's ``ResNet.restore``
(genData/network.py:113-122), whose only implementation is ``tf.train.Saver``.
A bundle is ``<prefix>.index`` -- a LevelDB table whose values are
``BundleEntryProto`` messages -- plus ``<prefix>.data-00000-of-00001`` holding
the raw little-endian tensors.  Only what the shipped checkpoints use is
supported: uncompressed blocks, one data shard, DT_FLOAT tensors.
"""
from __future__ import annotations

import os
import re
import struct

import numpy as np

_TABLE_MAGIC = 0xDB4775248B80FB57


def _varint(buf: bytes, pos: int):
    val = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if b < 0x80:
            return val, pos
        shift += 7


def _block_entries(raw: bytes, offset: int, size: int):
    """Yield (key, value) from one prefix-compressed table block."""
    if raw[offset + size] != 0:
        raise ValueError("compressed checkpoint index blocks are not supported")
    blk = raw[offset:offset + size]
    n_restarts = struct.unpack_from("<I", blk, len(blk) - 4)[0]
    end = len(blk) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(blk, pos)
        fresh, pos = _varint(blk, pos)
        vlen, pos = _varint(blk, pos)
        key = key[:shared] + blk[pos:pos + fresh]
        pos += fresh
        yield key, blk[pos:pos + vlen]
        pos += vlen


def _handle(buf: bytes, pos: int = 0):
    off, pos = _varint(buf, pos)
    size, pos = _varint(buf, pos)
    return off, size, pos


def _proto_fields(buf: bytes):
    """Minimal protobuf wire decoder: yields (field number, wire type, value)."""
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield num, wt, val


def _entry(buf: bytes):
    """BundleEntryProto -> (dtype enum, shape, shard, offset, size)."""
    dtype, shape, shard, offset, size = 0, [], 0, 0, 0
    for num, _, val in _proto_fields(buf):
        if num == 1:
            dtype = val
        elif num == 2:                       # TensorShapeProto { repeated Dim dim = 2 }
            for n2, _, dim in _proto_fields(val):
                if n2 == 2:
                    d = 0
                    for n3, _, v3 in _proto_fields(dim):
                        if n3 == 1:
                            d = v3
                    shape.append(d)
        elif num == 3:
            shard = val
        elif num == 4:
            offset = val
        elif num == 5:
            size = val
    return dtype, tuple(shape), shard, offset, size


def resolve_prefix(path: str) -> str:
    """A directory resolves through its ``checkpoint`` text file to the latest
    prefix (tf.train.get_checkpoint_state); otherwise ``path`` is the prefix itself.
    Raises FileNotFoundError like network.py:118-122."""
    if os.path.isdir(path):
        marker = os.path.join(path, "checkpoint")
        if os.path.exists(marker):
            m = re.search(r'model_checkpoint_path:\s*"([^"]+)"', open(marker).read())
            if m:
                name = m.group(1)
                path = name if os.path.isabs(name) else os.path.join(path, name)
    if not os.path.exists(path + ".index"):
        raise FileNotFoundError("Could not find old network weights")
    return path


def read_bundle(path: str) -> dict[str, np.ndarray]:
    """All variables of the checkpoint at ``path`` (directory or prefix) as float32
    arrays in TF layout (conv kernels HWIO, dense kernels [in, out])."""
    prefix = resolve_prefix(path)
    raw = open(prefix + ".index", "rb").read()
    footer = raw[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != _TABLE_MAGIC:
        raise ValueError("not a TensorFlow bundle index (bad table magic)")
    _, _, pos = _handle(footer, 0)                       # metaindex handle (unused)
    idx_off, idx_size, _ = _handle(footer, pos)
    data = np.memmap(prefix + ".data-00000-of-00001", dtype=np.uint8, mode="r")
    out = {}
    for _, hval in _block_entries(raw, idx_off, idx_size):
        off, size, _ = _handle(hval)
        for key, val in _block_entries(raw, off, size):
            if not key:                                   # BundleHeaderProto
                continue
            dtype, shape, shard, toff, tsize = _entry(val)
            if dtype != 1 or shard != 0:
                raise ValueError(f"{key!r}: only DT_FLOAT tensors in shard 0 are supported")
            arr = np.frombuffer(bytes(data[toff:toff + tsize]), dtype="<f4").reshape(shape)
            out[key.decode()] = arr.astype(np.float32)
    return out
