"""One-off GPU diagnostics (round 2): batch-position invariance of the device net; engine(N=1, host pv_fn = device net)
vs the oracle player on the roots of test_matches_oracle_on_many_roots_with_device_net."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden
from oracle import mcts as omcts, rules as orules
from alphafive_b200.net import DeviceNet, glorot_init
from alphafive_b200.engine import SearchEngine, make_config

S, sims = 11, 64
g = golden("replay_sample.npz")
boards, last = g["boards"][:48], g["last_action"][:48]
w = glorot_init(S, 3)
net = DeviceNet(S, 48, w)
x = np.stack([orules.input_planes(b, tuple(la) if la[0] >= 0 else None) for b, la in zip(boards, last)])
pb, vb = net.eval(x)
worst = 0
for j in range(48):
    p1, v1 = net.eval(x[j:j + 1])
    d = max(np.abs(p1[0] - pb[j]).max(), abs(v1[0] - vb[j]))
    worst = max(worst, d)
    if d > 0:
        print("board", j, "batch-position dependent:", d, "bitwise equal p:", (p1[0] == pb[j]).mean())
print("net batch invariance: worst abs diff", worst)
# shifted batch: same boards at other batch positions
xs = np.roll(x, 7, axis=0)
ps, vs = net.eval(xs)
print("rolled batch equal:", (np.roll(ps, -7, axis=0) == pb).all(), (np.roll(vs, -7, axis=0) == vb).all())

cells = np.where(last[:, 0] >= 0, last[:, 0].astype(np.int64) * S + last[:, 1], -1).astype(np.int32)
eng = SearchEngine(make_config(board_size=S, simulation_per_step=sims, upper_simulation_per_step=sims + 100, n_games=48, training=False))
eng.set_roots(boards, cells)
eng.run_search(net=net, check_every=4)
n_dev = eng.root_stats()[0].cpu().numpy()
eng1 = SearchEngine(make_config(board_size=S, simulation_per_step=sims, upper_simulation_per_step=sims + 100, n_games=1, training=False))
for j in range(48):
    la = tuple(int(v) for v in last[j]); la = la if la[0] >= 0 else None
    cfg = omcts.SearchConfig(simulation_per_step=sims, upper_simulation_per_step=sims + 100)
    pl = omcts.OraclePlayer(cfg, training=False, pv_fn=net.eval)
    pl.get_action(boards[j], la)
    n_or = pl.root_stats(boards[j])[0]
    eng1.set_roots(boards[j][None], cells[j:j + 1], None, np.ones(1, np.uint8))
    eng1.run_search(pv_fn=net.eval)
    n_host = eng1.root_stats()[0].cpu().numpy()[0]
    if not (n_or == n_dev[j]).all() or not (n_host == n_or).all():
        print("root", j, "oracle==lockstep-device-net:", (n_or == n_dev[j]).all(), " oracle==engine(host pv_fn=net.eval):", (n_host == n_or).all(),
              "diff cells", np.flatnonzero(n_or != n_dev[j])[:8], n_or[n_or != n_dev[j]][:8], n_dev[j][n_or != n_dev[j]][:8])
print("done")
