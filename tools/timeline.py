"""clock64 timeline of CTA 0 of every conv layer (GPU tooling).  python tools/timeline.py [layer 1..10] [groups]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from alphafive_b200 import _lib
from alphafive_b200._lib import check, ptr, stream_ptr
from alphafive_b200.net import DeviceNet, glorot_init
layers = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1]
layer = layers[0]
ngr = int(sys.argv[2]) if len(sys.argv) > 2 else 4
S, n = 11, 4096
lib = _lib.load()
fn = lib.a5__debug_timeline
fn.restype = C.c_int
fn.argtypes = [C.c_void_p] * 6 + [C.c_void_p]
net = DeviceNet(S, n, glorot_init(S, 0), mode=_lib.NET_TC)
planes = torch.from_numpy((np.random.default_rng(0).random((n, 3, S, S)) < 0.2).astype(np.int8)).cuda()
prob = torch.empty((n, S * S), device="cuda"); val = torch.empty((n,), device="cuda")
net.forward(planes, prob, val)
dbg = torch.zeros((10, 8, 256), dtype=torch.int64, device="cuda")
check(fn(net.handle, ptr(planes), n, ptr(prob), ptr(val), ptr(dbg), stream_ptr()))
torch.cuda.synchronize()
if layer == 0:      # summary: CTA-0 cycles per layer (clock-independent)
    tot = 0
    for l in range(10):
        dl = dbg.cpu().numpy()[l]
        nz = dl[dl > 0]
        span = int(nz.max() - nz.min())
        tot += span
        print(f"layer {l+1:2d}: {span:9d} ns")
    print(f"sum      : {tot:9d} ns")
    sys.exit(0)
for layer in layers:
    print(f"=== layer {layer}")
    d = dbg.cpu().numpy()[layer - 1]
    t0 = min(x for x in d.ravel() if x > 0)
    ev = []
    names = {0: "P slab-issue", 1: "M", 2: "E"}
    for role in range(8):
        for i, x in enumerate(d[role]):
            if x > 0:
                ev.append((int(x - t0), role, i))
    ev.sort()
    nslab = {1: 1, 2: 3, 3: 2, 4: 6, 5: 4, 6: 7, 7: 4, 8: 6, 9: 2, 10: 3}[layer]
    per_group_m = 3 + nslab
    for t, role, i in ev[: ngr * (per_group_m + 2 + nslab + 48 * nslab)]:
        if role == 1:
            k = i % per_group_m
            what = ["wait-tmem", "tmem-free"][k] if k < 2 else ("issued-all" if k == per_group_m - 1 else f"slab{k-2}-landed")
            print(f"{t:9d}  MMA g{i // per_group_m} {what}")
        elif role == 0:
            print(f"{t:9d}  PRODUCER slab {i} issue")
        elif role == 2:
            print(f"{t:9d}  EPI g{i // 2} {'ready' if i % 2 == 0 else 'drained'}")
        elif role == 3:
            print(f"{t:9d}      stage {i // 2} {'wait' if i % 2 == 0 else 'landed'}")
        elif role == 4:
            print(f"{t:9d}          W-leader issue stage {i}")
        elif role == 5:
            print(f"{t:9d}          RELAY stage {i} landed in peer")
        elif role == 6:
            print(f"{t:9d}          W-peer issue stage {i}")
