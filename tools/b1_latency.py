"""B = 1 latency of one search pass (net forward of one board + tree pass) under CUDA-graph replay (GPU tooling).
python tools/b1_latency.py [n_boards]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from alphafive_b200.engine import SearchEngine, make_config
from alphafive_b200.net import DeviceNet, glorot_init

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 2            # 1 tensor-core path, 2 the one-kernel latency path
S = 11
eng = SearchEngine(make_config(board_size=S, simulation_per_step=500, upper_simulation_per_step=600, n_games=n, training=False))
net = DeviceNet(S, n, glorot_init(S, 0))
prob = torch.zeros((n, S * S), device="cuda"); val = torch.zeros((n,), device="cuda")
eng.set_roots(np.zeros((n, S, S), np.int8), np.full(n, -1, np.int32), None, np.ones(n, np.uint8))
eng.step()


def timed(fn, reps=500):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        fn()
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps, (time.perf_counter() - t0) * 1e6 / reps


def both():
    net.forward_raw(eng.planes_ptr, n, prob, val, mode); eng.step(prob, val)


print(f"n = {n} mode {mode}: pass (forward + tree)  gpu {timed(both)[0]:7.1f} us")
print(f"        forward only           gpu {timed(lambda: net.forward_raw(eng.planes_ptr, n, prob, val, mode))[0]:7.1f} us")
print(f"        tree pass only         gpu {timed(lambda: eng.step(prob, val))[0]:7.1f} us")
t0 = time.perf_counter()
for _ in range(200):
    both()
torch.cuda.synchronize()
print(f"        eager (ctypes launches) wall {(time.perf_counter() - t0) * 1e6 / 200:7.1f} us per pass")
if mode == 2:
    import ctypes as C
    from alphafive_b200 import _lib
    lib = _lib.load()
    lib.a5__debug_small_timeline.argtypes = [C.c_void_p]
    dbg = torch.zeros(12, dtype=torch.int64, device="cuda")
    lib.a5__debug_small_timeline(C.c_void_p(dbg.data_ptr()))
    acc = np.zeros(11)
    for _ in range(50):
        net.forward_raw(eng.planes_ptr, n, prob, val, mode); torch.cuda.synchronize()
        d = dbg.cpu().numpy().astype(np.float64)
        acc += np.diff(d)
    lib.a5__debug_small_timeline(None)
    names = ["conv1", "b1c1", "b1c2", "b2c1", "b2c2", "b3c1+b4c1", "b3c2+b4c2", "b5c1+vconv", "b5c2+vfc1", "pconv+pfc", "final"]
    print("        phases (us, CTA 0, incl. the barrier): " + "  ".join(f"{nm} {a / 50 / 1e3:.1f}" for nm, a in zip(names, acc)))
