"""Run a few tensor-core forwards (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from alphafive_b200 import _lib
from alphafive_b200.net import DeviceNet, glorot_init
S = int(sys.argv[1]) if len(sys.argv) > 1 else 11
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
net = DeviceNet(S, n, glorot_init(S, 0), mode=_lib.NET_TC)
rng = np.random.default_rng(0)
planes = torch.from_numpy((rng.random((n, 3, S, S)) < 0.2).astype(np.int8)).cuda()
for _ in range(reps):
    p, v = net.forward(planes)
torch.cuda.synchronize()
print(float(p.sum()), float(v.sum()))
