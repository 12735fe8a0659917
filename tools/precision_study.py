"""CPU study: how many of the three split-precision MMA passes does the 1e-4 parity bar need?
Emulates rounding the block-conv operands to fp16 (11 significant bits) on the trained checkpoint and on
glorot weights and compares with a float64 evaluation.   python tools/precision_study.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from oracle import net as onet, rules as orules

def q16(t):           # round to fp16 precision, scale-free (operands are pre-scaled into the normal range)
    m, e = torch.frexp(t.double())
    return torch.ldexp(torch.round(m * 2048.0) / 2048.0, e)

class QNet(onet.OracleNet):
    def __init__(self, size, w, qa=(), qw=()):
        super().__init__(size, w, dtype=torch.float64)
        self.qa, self.qw = set(qa), set(qw)
    def _conv(self, x, name, act):
        k = self.k[name + "/kernel"]
        if name in self.qw: k = q16(k)
        if name in self.qa: x = q16(x)
        y = F.conv2d(x, k, self.w[name + "/bias"], padding=k.shape[-1] // 2)
        return F.elu(y) if act else y

def block_convs():
    out = []
    for name, _, _ in onet.BLOCKS:
        out += [name + "_res", name + "_conv1", name + "_conv2"]
    return out

S = 11
rng = np.random.default_rng(0)
boards = [orules.random_board(rng, S, d) for d in np.linspace(0.02, 0.6, 192)]
x = np.stack([orules.input_planes(b) for b in boards]).astype(np.float64)
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ckpt6960.npz"))
trained = {k.replace("__", "/"): g[k] for k in g.files}
for label, w in (("ckpt-6960", trained), ("glorot seed 0", onet.glorot_weights(S, 0))):
    if not w: continue
    ref_p, ref_v = QNet(S, w).eval(x)
    p32, v32 = onet.OracleNet(S, w).eval(x.astype(np.float32))
    print(f"== {label}: fp32 torch vs float64: dp {np.abs(p32 - ref_p).max():.2e} dv {np.abs(v32 - ref_v).max():.2e}")
    allc = block_convs()
    cases = {"weights fp16, activations exact (2 passes: a_hi w_hi + a_lo w_hi)": ((), allc),
             "activations fp16, weights exact (2 passes: a_hi w_hi + a_hi w_lo)": (allc, ()),
             "both fp16 (1 pass)": (allc, allc)}
    for name in allc:
        cases[f"only {name}: weights fp16"] = ((), [name])
    for cname, (qa, qw) in cases.items():
        p, v = QNet(S, w, qa, qw).eval(x)
        print(f"  {cname:70s} dp {np.abs(p - ref_p).max():.2e}  dv {np.abs(v - ref_v).max():.2e}")
