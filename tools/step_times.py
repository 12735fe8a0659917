"""Per-game time of one k_step in steady-state self-play (GPU tooling).  python tools/step_times.py [games] [sims]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from alphafive_b200 import _lib
from alphafive_b200._lib import check
from alphafive_b200.net import DeviceNet, glorot_init
from alphafive_b200.selfplay import SelfPlay
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 500
lib = _lib.load()
fn = lib.a5__debug_step_times
fn.restype = C.c_int
fn.argtypes = [C.c_int, C.c_void_p]
net = DeviceNet(11, N, glorot_init(11, 0), mode=_lib.NET_TC)
sp = SelfPlay(None, n_games=N, net=net, training=True, seed=0, use_graph=False, board_size=11,
              simulation_per_step=sims, upper_simulation_per_step=sims + 142)
sp.start()
sp.set_budget(40, 50); sp.run_passes(48 * 40); sp.harvest(); sp.set_budget(sims, sims + 142)
sp.run_passes(52)
g = torch.Generator(device="cuda"); g.manual_seed(7)
sl = sp.engine.sims_left()
sl.copy_(torch.minimum(sl, torch.randint(1, sims + 1, sl.shape, device=sl.device, dtype=torch.int32, generator=g)))   # spread the phases (bench.desync_budgets)
sp.run_passes(sims + 137)
check(fn(1, None))
buf = np.zeros((2, 8192), np.uint64)
tot = []
for rep in range(int(sys.argv[3]) if len(sys.argv) > 3 else 8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    net.forward_raw(sp.engine.planes_ptr, N, sp.prob, sp.value)
    e0.record(); sp.engine.step(sp.prob, sp.value); e1.record()
    torch.cuda.synchronize()
    check(fn(1, buf.ctypes.data))
    ns, fl = buf[0, :N].astype(np.int64), buf[1, :N].astype(np.int64)
    moved, fin, depth = (fl & 2) != 0, (fl & 4) != 0, fl >> 8
    print(f"pass {rep}: kernel {e0.elapsed_time(e1)*1e3:.1f} us | warp ns: mean {ns.mean():.0f} p50 {np.percentile(ns,50):.0f} "
          f"p99 {np.percentile(ns,99):.0f} max {ns.max()} | moved {moved.sum()} (mean {ns[moved].mean() if moved.any() else 0:.0f}, "
          f"max {ns[moved].max() if moved.any() else 0}) finished {fin.sum()} | not moved: mean {ns[~moved].mean():.0f} max {ns[~moved].max()} "
          f"| depth mean {depth.mean():.2f} max {depth.max()}")
    for d in range(0, 8):
        m = (depth == d) & ~moved
        if m.any():
            print(f"     depth {d}: {m.sum():5d} warps, mean {ns[m].mean():.0f} ns, max {ns[m].max()}")
check(fn(0, None))
