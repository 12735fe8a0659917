"""Feasibility probe: SM-partitioned streams (CUDA green contexts) for the two-half-batch pipeline
(DESIGN.md section 8, item 0).  python tools/green_probe.py [small_sms]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cuda.bindings import driver as cu
from alphafive_b200 import _lib
from alphafive_b200._lib import check, ptr, stream_ptr
from alphafive_b200.engine import SearchEngine, make_config
from alphafive_b200.net import DeviceNet, glorot_init

def ck(r):
    if r[0] != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(str(r[0]))
    return r[1:] if len(r) > 2 else (r[1] if len(r) == 2 else None)

small_sms = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.zeros(1).cuda()
dev = ck(cu.cuDeviceGet(0))
res = ck(cu.cuDeviceGetDevResource(dev, cu.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
print("device SMs", res.sm.smCount)
groups, nb, rem = ck(cu.cuDevSmResourceSplitByCount(1, res, 0, small_sms))
print("split: group", groups[0].sm.smCount, "remaining", rem.sm.smCount, "nb", nb)
streams = []
for r in (groups[0], rem):
    desc = ck(cu.cuDevResourceGenerateDesc([r], 1))
    g = ck(cu.cuGreenCtxCreate(desc, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
    st = ck(cu.cuGreenCtxStreamCreate(g, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
    streams.append(torch.cuda.ExternalStream(int(st)))
s_small, s_big = streams
n_small, n_big = groups[0].sm.smCount, rem.sm.smCount

S, N = 11, 2048
lib = _lib.load()
w = glorot_init(S, 0)
net = DeviceNet(S, N, w)
rng = np.random.default_rng(0)
planes = torch.from_numpy((rng.random((N, 3, S, S)) < 0.2).astype(np.int8)).cuda()
prob0, val0 = torch.empty((N, S * S), device="cuda"), torch.empty(N, device="cuda")
prob1, val1 = torch.empty_like(prob0), torch.empty_like(val0)
net.forward(planes, prob0, val0)
torch.cuda.synchronize()
parts = lambda p, pr, vl: check(lib.a5_net_forward_parts(net.handle, ptr(planes), N, ptr(pr), ptr(vl), p, stream_ptr()))

def timed(fn, stream, reps=10):
    with torch.cuda.stream(stream):
        fn(); fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

cur = torch.cuda.current_stream()
print(f"full device, n={N}: front {timed(lambda: parts(1, prob1, val1), cur):.1f} us, body {timed(lambda: parts(2, prob1, val1), cur):.1f} us, "
      f"heads {timed(lambda: parts(4, prob1, val1), cur):.1f} us")
check(lib.a5_net_set_sm_limit(net.handle, n_big, n_small))
with torch.cuda.stream(s_small):
    parts(1, prob1, val1)
torch.cuda.synchronize()
with torch.cuda.stream(s_big):
    parts(2, prob1, val1)
torch.cuda.synchronize()
with torch.cuda.stream(s_small):
    parts(4, prob1, val1)
torch.cuda.synchronize()
print("partitioned result equals full-device result:", bool((prob1 == prob0).all()), bool((val1 == val0).all()),
      float((prob1 - prob0).abs().max()))
t_body = timed(lambda: parts(2, prob1, val1), s_big)
t_front = timed(lambda: parts(1, prob1, val1), s_small)
t_heads = timed(lambda: parts(4, prob1, val1), s_small)
print(f"body on {n_big} SMs: {t_body:.1f} us; front on {n_small} SMs: {t_front:.1f} us; heads on {n_small} SMs: {t_heads:.1f} us")

eng = SearchEngine(make_config(board_size=S, simulation_per_step=500, upper_simulation_per_step=642, n_games=N,
                               training=True, auto_play=True))
with torch.cuda.stream(s_small):
    eng.step()
    for _ in range(30):
        parts(7, prob1, val1) if False else net.forward_raw(eng.planes_ptr, N, prob1, val1)
        eng.step(prob1, val1)
torch.cuda.synchronize()
t_step = timed(lambda: eng.step(prob1, val1), s_small)
print(f"tree pass for {N} games on {n_small} SMs: {t_step:.1f} us")

# concurrency: body on the big partition while front + heads + tree pass run on the small one
def both(reps=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(s_big):
        e0.record()
        for _ in range(reps):
            parts(2, prob1, val1)
        e1.record()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s_small):
        f0.record()
        for _ in range(reps):
            parts(4, prob0, val0); eng.step(prob0, val0); parts(1, prob0, val0)
        f1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, f0.elapsed_time(f1) / reps * 1e3
both(3)
b, s = both(10)
print(f"concurrent: body {b:.1f} us per iteration on the big partition, heads+tree+front {s:.1f} us on the small one")
