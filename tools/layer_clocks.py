"""Per-layer duration, SM clock and CTA-0 cycles of the block-conv launches inside a continuous stream of forwards
(GPU tooling; a5__debug_clk stamps clock64 + %globaltimer at start / end of CTA 0; the time of a launch includes its
programmatic-dependent-launch wait for the predecessor).  python tools/layer_clocks.py [reps] [boards per forward]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from alphafive_b200 import _lib
from alphafive_b200._lib import check, ptr, stream_ptr
from alphafive_b200.net import DeviceNet, glorot_init
S = 11
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
lib = _lib.load()
lib.a5__debug_clk.argtypes = [C.c_void_p]; lib.a5__debug_clk.restype = C.c_int
net = DeviceNet(S, n, glorot_init(S, 0), mode=_lib.NET_TC)
planes = torch.from_numpy((np.random.default_rng(0).random((n, 3, S, S)) < 0.2).astype(np.int8)).cuda()
prob = torch.empty((n, S * S), device="cuda"); val = torch.empty((n,), device="cuda")
clk = torch.zeros((8, 4), dtype=torch.int64, device="cuda")
check(lib.a5__debug_clk(ptr(clk)))
acc = np.zeros((8, 2))
for r in range(reps):
    net.forward(planes, prob, val)
    if r >= reps // 2 and r % 10 == 0:
        torch.cuda.synchronize()
        c = clk.cpu().numpy().astype(np.float64)
        acc[:, 0] += c[:, 2] - c[:, 0]; acc[:, 1] += c[:, 3] - c[:, 1]
torch.cuda.synchronize()
names = ["b1c1", "b1c2", "b2c1", "b2c2", "mc1", "mc2", "b5c1", "b5c2"]
k = len(range(reps // 2, reps, 10))
for i, nm in enumerate(names):
    print(f"{nm:5s} {acc[i,1]/k/1e3:8.1f} us  {acc[i,0]/acc[i,1]*1e3:7.1f} MHz  {acc[i,0]/k:10.0f} cycles")
print(f"sum   {acc[:,1].sum()/k/1e3:8.1f} us for {n} boards  ({acc[:,1].sum()/k/1e3*4096/n:8.1f} us per 4096)  mean clock {acc[:,0].sum()/acc[:,1].sum()*1e3:7.1f} MHz")
