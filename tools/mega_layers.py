"""Per-layer time and SM clock inside the chunk-major megakernel (A5_TC_MEGA=1), steady state (GPU tooling)."""
import ctypes as C, os, sys
os.environ["A5_TC_MEGA"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from alphafive_b200 import _lib
from alphafive_b200._lib import check, ptr
from alphafive_b200.net import DeviceNet, glorot_init
S, n = 11, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = 400
lib = _lib.load()
lib.a5__debug_mega_clk.argtypes = [C.c_void_p]
net = DeviceNet(S, n, glorot_init(S, 0))
planes = torch.from_numpy((np.random.default_rng(0).random((n, 3, S, S)) < 0.2).astype(np.int8)).cuda()
prob = torch.empty((n, S * S), device="cuda"); val = torch.empty((n,), device="cuda")
nl, nch = 8, 4
clk = torch.zeros((nch * nl + 1, 2), dtype=torch.int64, device="cuda")
check(lib.a5__debug_mega_clk(ptr(clk)))
acc = np.zeros((nl, 2)); tot = np.zeros(2); k = 0
for r in range(reps):
    net.forward(planes, prob, val)
    if r >= reps // 2 and r % 10 == 0:
        torch.cuda.synchronize()
        c = clk.cpu().numpy().astype(np.float64)
        d = np.diff(c, axis=0).reshape(nch, nl, 2)
        acc += d.sum(0); tot += c[-1] - c[0]; k += 1
names = ["b1c1", "b1c2", "b2c1", "b2c2", "mc1", "mc2", "b5c1", "b5c2"]
for i, nm in enumerate(names):
    print(f"{nm:5s} {acc[i,0]/k/1e3:8.1f} us  {acc[i,1]/acc[i,0]*1e3:7.1f} MHz  {acc[i,1]/k:10.0f} cycles")
print(f"kernel {tot[0]/k/1e3:8.1f} us  mean clock {tot[1]/tot[0]*1e3:7.1f} MHz  {tot[1]/k:10.0f} cycles  (issue-side stamps of CTA 0)")
