"""Device net (tensor-core path) against the fp32 oracle on the 1000 replay boards, ckpt-6960 and glorot weights
(GPU tooling; used with A5_TC_LO_DROP / A5_TC_HI_DROP to map the operand-precision / clock trade, profiles/r02_lo_bits.txt)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
from conftest import golden
from oracle import net as onet, rules as orules
from alphafive_b200.net import DeviceNet, glorot_init
g = golden("replay_sample.npz")
x = np.stack([orules.input_planes(b, tuple(la) if la[0] >= 0 else None) for b, la in zip(g["boards"], g["last_action"])])
z = golden("ckpt6960.npz"); w = {k.replace("__", "/"): z[k] for k in z.files}
for name, ww in (("ckpt6960", w), ("glorot", glorot_init(11, 0))):
    wp, wv = onet.OracleNet(11, ww).eval(x)
    net = DeviceNet(11, 1024, ww)
    gp, gv = net.eval(x)
    print(f"{name}: max|dp| {np.abs(gp-wp).max():.3e}  max|dv| {np.abs(gv-wv).max():.3e}  rms dv {np.sqrt(((gv-wv)**2).mean()):.2e}  argmax agree {(gp.argmax(1)==wp.argmax(1)).mean():.4f}")
