"""A5_TC_MEGA: the chunk-major megakernel against the layer-per-launch path (bitwise) and its forward time (GPU tooling)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from alphafive_b200.net import DeviceNet, glorot_init

def make(S, n, mega):
    os.environ["A5_TC_MEGA"] = "1" if mega else "0"
    return DeviceNet(S, n, glorot_init(S, 0))

def timed(net, planes, prob, val, reps=60):
    for _ in range(10):
        net.forward(planes, prob, val)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        net.forward(planes, prob, val)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps

for S, sizes in ((11, (1, 2, 48, 300, 1024, 4096)), (15, (3, 200, 2048)), (9, (77,)), (13, (500,))):
    for n in sizes:
        rng = np.random.default_rng(n)
        planes = torch.from_numpy((rng.random((n, 3, S, S)) < 0.25).astype(np.int8)).cuda()
        a, b = make(S, n, False), make(S, n, True)
        pa, va = a.forward(planes); pb, vb = b.forward(planes)
        torch.cuda.synchronize()
        same = torch.equal(pa, pb) and torch.equal(va, vb)
        print(f"S={S} n={n}: bitwise equal {same}  max|dp| {float((pa-pb).abs().max()):.2e}", flush=True)
        if n >= 1024:
            prob = torch.empty_like(pa); val = torch.empty_like(va)
            ta, tb = timed(a, planes, prob, val), timed(b, planes, prob, val)
            ta2, tb2 = timed(a, planes, prob, val), timed(b, planes, prob, val)
            print(f"        forward: layer-per-launch {ta:.1f} / {ta2:.1f} us   megakernel {tb:.1f} / {tb2:.1f} us", flush=True)
        a.close(); b.close()
