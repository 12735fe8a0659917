"""TensorFlow checkpoint-v2 ("bundle") reader and writer, without TensorFlow.

Replaces the reference's ``ResNet.restore`` (genData/network.py:113-122) and the
``net.saver.save`` calls of its trainer (main.py:74-77), whose only implementation is
``tf.train.Saver``: ``read_bundle`` loads the shipped checkpoints, ``write_bundle`` writes
files the reference's ``restore`` can load (for the 42 variables of ckpt/alphaFive-6960 it
reproduces the shipped ``.index`` and ``.data`` files byte for byte -- tests/test_ckpt.py).
A bundle is ``<prefix>.index`` -- a LevelDB table whose values are
``BundleEntryProto`` messages -- plus ``<prefix>.data-00000-of-00001`` holding
the raw little-endian tensors.  Only what the shipped checkpoints use is
supported: uncompressed blocks, one data shard, DT_FLOAT tensors.
"""
from __future__ import annotations

import os
import re
import struct
import warnings

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
                # e.g. an int64 global_step: load_pretrained of the reference skips what it cannot use
                # (network.py:136-160); so does this reader, but it says so
                warnings.warn(f"checkpoint entry {key.decode()!r} skipped (dtype {dtype}, shard {shard}): "
                              "only DT_FLOAT tensors of shard 0 are read")
                continue
            arr = np.frombuffer(bytes(data[toff:toff + tsize]), dtype="<f4").reshape(shape)
            out[key.decode()] = arr.astype(np.float32)
    return out


# ---------------------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------------------
def _crc32c_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TAB = None
_CRC_LANES = 2048          # the data is cut into this many equal chunks that advance in lock-step (numpy)


def _crc_tab():
    global _CRC_TAB
    if _CRC_TAB is None:
        _CRC_TAB = np.array(_crc32c_table(), np.uint32)
    return _CRC_TAB


def _zero_op(nbytes: int):
    """GF(2) matrix (32 column ints) of "feed `nbytes` zero bytes" on the raw CRC register."""
    tab = _crc_tab()
    one = [int(tab[(1 << j) & 0xFF]) ^ ((1 << j) >> 8) for j in range(32)]      # one zero byte

    def apply(mat, v):
        r, j = 0, 0
        while v:
            if v & 1:
                r ^= mat[j]
            v >>= 1
            j += 1
        return r

    result = [1 << j for j in range(32)]
    sq = one
    while nbytes:
        if nbytes & 1:
            result = [apply(sq, c) for c in result]
        sq = [apply(sq, c) for c in sq]
        nbytes >>= 1
    return result, apply


def crc32c(data: bytes) -> int:
    """CRC-32C (Castagnoli), the checksum of LevelDB tables and bundle entries.  The register update is
    linear over GF(2): the data (zero-padded at the front, which leaves a zero register unchanged) is cut
    into _CRC_LANES chunks whose registers advance together as numpy vectors, the chunk registers are folded
    with the "append m zero bytes" operator, and the initial 0xFFFFFFFF is carried through the same way."""
    n = len(data)
    tab = _crc_tab()
    if n < 4 * _CRC_LANES:
        c = 0xFFFFFFFF
        for b in data:
            c = int(tab[(c ^ b) & 0xFF]) ^ (c >> 8)
        return c ^ 0xFFFFFFFF
    m = -(-n // _CRC_LANES)
    buf = np.zeros(_CRC_LANES * m, np.uint8)
    buf[_CRC_LANES * m - n:] = np.frombuffer(data, np.uint8)
    cols = buf.reshape(_CRC_LANES, m).T.astype(np.uint32)          # cols[t] = byte t of every chunk
    reg = np.zeros(_CRC_LANES, np.uint32)
    for t in range(m):
        reg = tab[(reg ^ cols[t]) & 0xFF] ^ (reg >> 8)
    zm, apply = _zero_op(m)
    acc = 0
    for r in reg.tolist():
        acc = apply(zm, acc) ^ r
    zn, _ = _zero_op(n)
    return (acc ^ apply(zn, 0xFFFFFFFF)) ^ 0xFFFFFFFF


def _masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _build_block(entries, restart_interval=16) -> bytes:
    """LevelDB block: prefix-compressed (key, value) entries + restart array + restart count."""
    out, restarts, last, n = bytearray(), [], b"", 0
    for key, val in entries:
        shared = 0
        if n % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(last), len(key)) and last[shared] == key[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(val)) + key[shared:] + val
        last, n = key, n + 1
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _entry_proto(shape, offset: int, size: int, crc: int, dtype: int = 1) -> bytes:
    """BundleEntryProto: dtype (1 = DT_FLOAT), shape, (offset), size, crc32c (masked, fixed32)."""
    dims = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(n)) for n in shape))
    out = b"\x08" + _put_varint(dtype) + b"\x12" + _put_varint(len(dims)) + dims
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size) + b"\x35" + struct.pack("<I", crc)
    return out


def _short_successor(key: bytes) -> bytes:
    """leveldb BytewiseComparator::FindShortSuccessor: first byte that can be incremented, truncated there."""
    for i, b in enumerate(key):
        if b != 0xFF:
            return key[:i] + bytes([b + 1])
    return key


def write_bundle(prefix: str, tensors: dict, update_marker: bool = True) -> None:
    """Write ``<prefix>.index`` and ``<prefix>.data-00000-of-00001`` (and the directory's ``checkpoint``
    marker) for float32 ``tensors`` in TF layout, as ``tf.train.Saver.save`` lays a one-shard bundle out:
    tensors concatenated in key order, one table block of BundleEntryProto values (restart interval 16),
    an empty metaindex block, a one-entry index block, and the 48-byte footer."""
    names = sorted(tensors)
    data = bytearray()
    entries = [(b"", b"\x08\x01\x1a\x02\x08\x01")]              # BundleHeaderProto: num_shards 1, version.producer 1
    for name in names:
        arr = np.asarray(tensors[name])
        code = {"i8": 9, "i4": 3}.get(arr.dtype.str[1:], 1)          # DT_INT64 / DT_INT32 (e.g. a global_step), else DT_FLOAT
        arr = arr.astype("<f4") if code == 1 else arr.astype(arr.dtype.newbyteorder("<"))
        raw = arr.tobytes(order="C")
        entries.append((name.encode(), _entry_proto(arr.shape, len(data), len(raw), _masked_crc(raw), code)))
        data += raw

    def with_trailer(block: bytes) -> bytes:
        return block + b"\x00" + struct.pack("<I", _masked_crc(block + b"\x00"))

    data_block = _build_block(entries)
    meta_block = _build_block([])
    out = bytearray(with_trailer(data_block))
    meta_off = len(out)
    out += with_trailer(meta_block)
    index_block = _build_block([(_short_successor(names[-1].encode()), _put_varint(0) + _put_varint(len(data_block)))], 1)
    index_off = len(out)
    out += with_trailer(index_block)
    footer = _put_varint(meta_off) + _put_varint(len(meta_block)) + _put_varint(index_off) + _put_varint(len(index_block))
    out += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _TABLE_MAGIC)
    folder = os.path.dirname(os.path.abspath(prefix))
    os.makedirs(folder, exist_ok=True)

    def put(path, payload, mode):
        tmp = f"{path}.tmp{os.getpid()}"                      # readers never see a half-written file
        with open(tmp, mode) as f:
            f.write(payload)
        os.replace(tmp, path)

    put(prefix + ".data-00000-of-00001", bytes(data), "wb")
    put(prefix + ".index", bytes(out), "wb")                  # the index last: it is what restore() looks for
    if update_marker:
        base = os.path.basename(prefix)
        marker = os.path.join(folder, "checkpoint")
        older = []
        if os.path.exists(marker):                            # tf.train.Saver keeps the last max_to_keep = 5 paths
            older = [m for m in re.findall(r'all_model_checkpoint_paths:\s*"([^"]+)"', open(marker).read()) if m != base]
        paths = (older + [base])[-5:]
        put(marker, f'model_checkpoint_path: "{base}"\n' + "".join(f'all_model_checkpoint_paths: "{m}"\n' for m in paths), "w")
