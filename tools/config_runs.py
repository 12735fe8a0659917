"""The other BASELINE.json configurations, timed briefly (not bench lines; GPU tooling):
  cfg 4: 4096 concurrent 15x15 games, 800 sims/move (upper 900), lock-step self-play
  cfg 5: arena, 1024 paired games between ckpt-6960 and a glorot net, 400 sims/move (first plies)
    python tools/config_runs.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from alphafive_b200.net import DeviceNet, glorot_init
from alphafive_b200.selfplay import SelfPlay

def cfg4():
    S, N, sims, upper = 15, 4096, 800, 900
    net = DeviceNet(S, N, glorot_init(S, 0))
    sp = SelfPlay(None, n_games=N, net=net, training=True, seed=0, board_size=S, simulation_per_step=sims,
                  upper_simulation_per_step=upper)
    sp.start()
    sp.set_budget(40, 50); sp.run_passes(24 * 40); sp.harvest(); sp.set_budget(sims, upper)
    sp.run_passes(sims)
    c0 = sp.counters(); torch.cuda.synchronize(); t0 = time.perf_counter()
    sp.run_passes(sims); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    c1 = sp.counters()
    print(f"cfg4 15x15/800: {(c1['moves'] - c0['moves']) / dt:.0f} moves/s, {(c1['leaf_evals'] - c0['leaf_evals']) / dt / 1e6:.2f} M leaf evals/s, "
          f"{dt / sims * 1e3:.2f} ms per pass, overflows {c1['overflows']}")

def cfg5():
    import types
    from alphafive_b200 import config as base
    from alphafive_b200.drivers import Arena
    S, N = 11, 1024
    z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ckpt6960.npz"))
    w = {k.replace("__", "/"): z[k] for k in z.files}
    cfg = types.SimpleNamespace(**{k: v for k, v in vars(base).items() if not k.startswith("_")})
    cfg.board_size, cfg.simulation_per_step, cfg.upper_simulation_per_step = S, 400, 500
    a = Arena(cfg, DeviceNet(S, N, w), DeviceNet(S, N, glorot_init(S, 0)), N, seed=0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = a.play(max_plies=8)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"cfg5 arena 1024 games x 8 plies at 400 sims: {r['moves'] / dt:.0f} moves/s, {r['leaf_evals'] / dt / 1e6:.2f} M leaf evals/s "
          f"({dt:.1f} s); finished so far: {(r['winners'] >= 0).sum()}")

if __name__ == "__main__":
    cfg4()
    cfg5()
