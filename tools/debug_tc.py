"""Layer-by-layer comparison of the tensor-core path against the fp32 path (GPU tooling)."""
import ctypes as C
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from alphafive_b200 import _lib
from alphafive_b200.net import DeviceNet, glorot_init
from alphafive_b200._lib import check, ptr, stream_ptr

S = int(sys.argv[1]) if len(sys.argv) > 1 else 11
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
use_ckpt = len(sys.argv) > 3 and sys.argv[3] == "ckpt"
lib = _lib.load()
lib.a5__debug_activation.restype = C.c_int
lib.a5__debug_activation.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
if use_ckpt:
    z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ckpt6960.npz"))
    w = {k.replace("__", "/"): z[k] for k in z.files}
else:
    w = glorot_init(S, 0)
net = DeviceNet(S, n, w)
rng = np.random.default_rng(0)
planes = torch.from_numpy((rng.random((n, 3, S, S)) < 0.2).astype(np.int8)).cuda()
names = ["a32", "b1h", "b1o", "b2h", "b2o", "b3h", "b3o", "b4h", "b4o", "b5h", "b5o"]
chs = [32, 64, 64, 128, 128, 32, 32, 64, 64, 32, 32]
pb = (S + 1) ** 2
p0, v0 = net.forward(planes, mode=0)
torch.cuda.synchronize()
acts0 = []
for i, ch in enumerate(chs):
    out = torch.empty((n * pb, ch), dtype=torch.float32, device="cuda")
    check(lib.a5__debug_activation(net.handle, 0, i, n, ptr(out), stream_ptr()))
    acts0.append(out.clone())
lib.a5__debug_keep_head_acts(1)
p1, v1 = net.forward(planes, mode=1)
torch.cuda.synchronize()
print("tc forward done")
for i, ch in enumerate(chs):
    out = torch.empty((n * pb, ch), dtype=torch.float32, device="cuda")
    check(lib.a5__debug_activation(net.handle, 1, i, n, ptr(out), stream_ptr()))
    d = (out - acts0[i]).abs()
    print(f"{names[i]:4s} ch={ch:3d} max|ref|={acts0[i].abs().max().item():9.4f} max|diff|={d.max().item():.3e} "
          f"mean|diff|={d.mean().item():.3e} argmax={np.unravel_index(int(d.argmax()), d.shape)}")
print("prob diff", (p1 - p0).abs().max().item(), "value diff", (v1 - v0).abs().max().item())
for mode in (0, 1):
    for _ in range(3):
        net.forward(planes, mode=mode)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10):
        net.forward(planes, mode=mode)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 10
    print(f"mode {mode}: {dt*1e3:.3f} ms per forward of {n} boards -> {118.727e6*n/dt/1e12:.1f} TFLOP/s (11x11 flops)")
