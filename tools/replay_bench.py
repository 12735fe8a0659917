"""HBM roofline of the replay-sampling kernel (a5_replay_sample): GB/s of algorithmic bytes
(record read 32 + C + 4C, batch write 16C + 8 per sample) against the measured copy bandwidth.
    python tools/replay_bench.py [plies] """
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from alphafive_b200.utils import RandomStack

plies = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
S, C = 11, 121
st = RandomStack(S, length=plies)
st.ring.copy_(torch.randint(0, 255, st.ring.shape, dtype=torch.uint8, device=st.ring.device))
st.ring[:, 16:20] = 0                                            # last_action = 0
st.count = plies
peak = 6552.3
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
rng = np.random.default_rng(0)
for num in (512, 2048, 65536, 1 << 20):
    idx = rng.choice(plies, size=min(num, plies), replace=False)
    rot = rng.integers(0, 4, len(idx)).astype(np.uint8)
    flip = rng.integers(0, 2, len(idx)).astype(np.uint8)
    d_idx, d_rot, d_flip = (torch.from_numpy(x).cuda() for x in (idx.astype(np.int64), rot, flip))
    n = len(idx)
    out = [torch.empty((n, 3, S, S), device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, device="cuda"),
           torch.empty((n, C), device="cuda")]
    from alphafive_b200._lib import check, ptr, stream_ptr
    call = lambda: check(st.lib.a5_replay_sample(ptr(st.ring), S, ptr(d_idx), ptr(d_rot), ptr(d_flip), n, ptr(out[0]),
                                                 ptr(out[1]), ptr(out[2]), ptr(out[3]), stream_ptr()))
    for _ in range(3):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    byts = n * (32 + C + 4 * C + 16 * C + 8)
    print(f"batch {n:8d}: {us:9.1f} us  {n / us:8.2f} M samples/s  {byts / us / 1e3:8.1f} GB/s algorithmic "
          f"= {byts / us / 1e3 / peak:.3f} of the measured HBM copy peak ({peak:.0f} GB/s)")
