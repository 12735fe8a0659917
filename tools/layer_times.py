"""Per-launch-group timing of the tensor-core forward (GPU tooling; CUDA events between launches).

    python tools/layer_times.py [S] [n] [reps]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from alphafive_b200 import _lib
from alphafive_b200._lib import check, ptr, stream_ptr
from alphafive_b200.net import DeviceNet, glorot_init

S = int(sys.argv[1]) if len(sys.argv) > 1 else 11
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
lib = _lib.load()
fn = lib.a5__debug_layer_times
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_void_p]
net = DeviceNet(S, n, glorot_init(S, 0), mode=_lib.NET_TC)
rng = np.random.default_rng(0)
planes = torch.from_numpy((rng.random((n, 3, S, S)) < 0.2).astype(np.int8)).cuda()
prob = torch.empty((n, S * S), device="cuda")
val = torch.empty((n,), device="cuda")
ms = (C.c_float * 12)()
check(fn(net.handle, ptr(planes), n, 3, ptr(prob), ptr(val), ms, stream_ptr()))        # warm-up
check(fn(net.handle, ptr(planes), n, reps, ptr(prob), ptr(val), ms, stream_ptr()))
names = ["conv1", "b1c1 32>64", "b1c2 64>64+r32", "b2c1 64>128", "b2c2 128>128+r64", "b3c1+b4c1 128>96", "b3c2+b4c2 +r128>96",
         "(merged into b3c1)", "(merged into b3c2)", "b5c1 64>32", "b5c2 32>32+r64", "heads"]
kmac = [0, 288 * 64, 608 * 64, 576 * 128, 1216 * 128, 1152 * 96, 416 * 32 + 704 * 64, 0, 0, 576 * 32, 352 * 32, 0]
pos = n * (S + 1) ** 2
tot = 0.0
for i, nm in enumerate(names):
    tf = 2 * 3 * kmac[i] * pos / (ms[i] * 1e-3) / 1e12 if kmac[i] else 0.0
    print(f"{nm:18s} {ms[i]*1e3:8.1f} us   issued {tf:7.1f} TFLOP/s (3 passes, padded positions)")
    tot += ms[i]
print(f"total {tot*1e3:.1f} us -> {n/tot*1e3:.0f} leaf evals/s; algorithmic {118.727e6*(S*S/121)*n/(tot*1e-3)/1e12:.1f} TFLOP/s")
