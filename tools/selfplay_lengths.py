"""Game lengths / results of device self-play with ckpt-6960 (training mode, config.py budgets) against the lengths of the
470 games in the reference's shipped replay buffer (GPU tooling; the statistical pin of tests/test_gpu_stochastic.py)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from conftest import golden
from alphafive_b200.net import DeviceNet
from alphafive_b200.selfplay import SelfPlay
from alphafive_b200.engine import parse_records
z = golden("ckpt6960.npz"); w = {k.replace("__", "/"): z[k] for k in z.files}
N, sims, upper = 2048, 542, 642
net = DeviceNet(11, N, w)
sp = SelfPlay(None, n_games=N, net=net, training=True, seed=7, board_size=11, simulation_per_step=sims, upper_simulation_per_step=upper)
lens, res = [], []
want = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
while len(lens) < want:
    sp.run_passes(sims * 4)
    buf, _ = sp.harvest()
    recs = parse_records(buf, 11)
    seen = set()
    for r in recs:
        key = (r["game_id"], r["game_serial"])
        if key not in seen:
            seen.add(key); lens.append(r["game_len"]); res.append(r["result"])
lens, res = np.array(lens), np.array(res)
acc = 1.0 - np.clip(-0.0682 * lens + 1.364, 0.0, 1.0)          # utils.py:80-82 acceptance of a game of this length
print("games", len(lens), "raw mean len %.2f sd %.2f" % (lens.mean(), lens.std()), "black wins %.3f draws %.3f" % ((res == 1).mean(), (res == 0).mean()))
m = (acc * lens).sum() / acc.sum()
sd = np.sqrt((acc * (lens - m) ** 2).sum() / acc.sum())
print("acceptance-weighted mean len %.2f sd %.2f  (buffer: 25.53 / 8.49 over 470 games)" % (m, sd))
np.save(os.path.join(R, "gpurun_out", "selfplay_lens_6960.npy"), np.stack([lens, res]))
