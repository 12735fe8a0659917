"""Two half batches in a software pipeline on SM-partitioned streams.

One pass of the hot path is ``conv1 -> nine block convs -> heads -> tree pass`` and only the block
convs keep the tensor cores busy; heads, tree pass, input bitboards and conv1 are 11 % of the pass
during which they idle.  The block-conv CTAs fill every SM they run on (222 KB shared memory), so the
small kernels cannot share an SM with them; instead the GPU is split with CUDA green contexts into a
large partition (132 SMs on a B200) that runs nothing but block convs and a small one (16 SMs) for
everything else, and the batch into two halves that alternate::

    big   stream :  body(0)            | body(1)             | body(0) ...
    small stream :  ... chain(1)       | chain(0)            | chain(1) ...      chain = heads, tree pass,
                                                                                 bitboards + conv1

``body(h)`` waits for ``chain(h)`` of the previous tick (event), ``chain(h)`` for ``body(h)``.  Games are
independent and the network is evaluated per board, so results are bit-identical to the single-stream
schedule (tests/test_gpu_pipeline.py).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

PART_FRONT, PART_BODY, PART_HEADS = 1, 2, 4


class SmPartition:
    """``small`` / ``big``: torch streams bound to two disjoint SM sets (green contexts of the current
    device's primary context).  Raises RuntimeError when the driver cannot split the SMs."""

    _cache: dict = {}

    def __init__(self, small_sms: int = 16, device: int | None = None):
        from cuda.bindings import driver as cu
        self.device = torch.cuda.current_device() if device is None else device
        torch.zeros(1, device=f"cuda:{self.device}")                     # primary context is live

        def ck(r):
            if r[0] != cu.CUresult.CUDA_SUCCESS:
                raise RuntimeError(f"green context setup failed: {r[0]}")
            return r[1:] if len(r) > 2 else r[1]

        dev = ck(cu.cuDeviceGet(self.device))
        res = ck(cu.cuDeviceGetDevResource(dev, cu.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
        groups, _, rem = ck(cu.cuDevSmResourceSplitByCount(1, res, 0, small_sms))
        self.n_small, self.n_big = int(groups[0].sm.smCount), int(rem.sm.smCount)
        if self.n_big < 2 or self.n_small < 1:
            raise RuntimeError("SM split left an empty partition")
        self._ctx, streams = [], []
        for r in (groups[0], rem):
            desc = ck(cu.cuDevResourceGenerateDesc([r], 1))
            g = ck(cu.cuGreenCtxCreate(desc, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
            st = ck(cu.cuGreenCtxStreamCreate(g, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
            self._ctx.append((g, st))
            streams.append(torch.cuda.ExternalStream(int(st), device=f"cuda:{self.device}"))
        self.small, self.big = streams

    @classmethod
    def get(cls, small_sms: int = 16, device: int | None = None):
        key = (torch.cuda.current_device() if device is None else device, small_sms)
        if key not in cls._cache:
            cls._cache[key] = cls(small_sms, key[0])
        return cls._cache[key]


class TwoHalfPipeline:
    """Drives two (engine, net) halves.  Invariant between ticks: for each half the leaf planes of the
    next evaluation are staged and its FRONT part has been enqueued (event ``f[h]``)."""

    def __init__(self, engines, nets, part: SmPartition):
        assert len(engines) == 2 and len(nets) == 2
        self.eng, self.net, self.part = engines, nets, part
        self.lib = _lib.load()
        dev = engines[0].device
        self.prob = [torch.zeros((e.N, e.C), dtype=torch.float32, device=dev) for e in engines]
        self.value = [torch.zeros((e.N,), dtype=torch.float32, device=dev) for e in engines]
        for n in nets:
            check(self.lib.a5_net_set_sm_limit(n.handle, part.n_big, part.n_small))
        self.f = [torch.cuda.Event(), torch.cuda.Event()]
        self.b = [torch.cuda.Event(), torch.cuda.Event()]
        self.primed = False

    def _parts(self, h, parts):
        e = self.eng[h]
        check(self.lib.a5_net_forward_parts(self.net[h].handle, e.planes_ptr, e.N, ptr(self.prob[h]), ptr(self.value[h]),
                                            parts, stream_ptr()))

    def prime(self):
        """First descent of both halves (no network output yet) and their FRONT parts."""
        cur = torch.cuda.current_stream()
        self.part.small.wait_stream(cur)
        self.part.big.wait_stream(cur)
        with torch.cuda.stream(self.part.small):
            for h in (0, 1):
                self.eng[h].step()
                self._parts(h, PART_FRONT)
                self.f[h].record()
        self.primed = True

    def run(self, ticks: int):
        """``ticks`` passes: every game of both halves completes one simulation per tick."""
        assert self.primed
        small, big = self.part.small, self.part.big
        for _ in range(ticks):
            for h in (0, 1):
                with torch.cuda.stream(big):
                    big.wait_event(self.f[h])
                    self._parts(h, PART_BODY)
                    self.b[h].record()
                with torch.cuda.stream(small):
                    small.wait_event(self.b[h])
                    self._parts(h, PART_HEADS)
                    self.eng[h].step(self.prob[h], self.value[h])
                    self._parts(h, PART_FRONT)
                    self.f[h].record()

    def drain(self):
        """Make the current stream wait for everything enqueued on the two partitions."""
        cur = torch.cuda.current_stream()
        cur.wait_stream(self.part.small)
        cur.wait_stream(self.part.big)
