"""How many of the leaf positions a lock-step self-play batch sends to the network are duplicates -- within a pass, and of
leaves evaluated during the last K passes (GPU tooling: sizing of a cross-game evaluation cache)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from alphafive_b200.selfplay import SelfPlay
N, sims = 4096, 500
sp = SelfPlay(None, n_games=N, training=True, seed=0, board_size=11, simulation_per_step=sims, upper_simulation_per_step=642, use_graph=False)
sp.start()
sp.set_budget(40, 50); sp.run_passes(48 * 40); sp.harvest(); sp.set_budget(sims, 642)     # preroll as bench.py
sp.run_passes(1500)
w = torch.randint(1, 2**62, (3 * 121,), device="cuda", dtype=torch.int64)
tot = within = 0
hits = {k: 0 for k in (1, 8, 64, 512, 4096)}
hist = []
for p in range(600):
    sp.run_passes(1)
    need = sp.engine.need_eval().bool()
    x = sp.engine.planes().reshape(N, -1).to(torch.int64)
    h = (x * w).sum(1)[need]                      # 64-bit hash of the planes (wraps mod 2^64)
    u = torch.unique(h)
    tot += h.numel(); within += h.numel() - u.numel()
    us = set(u.cpu().tolist())
    for k in hits:
        seen = set().union(*hist[-k:]) if hist else set()
        hits[k] += len(us & seen)
    hist.append(us)
    if len(hist) > 4096: hist.pop(0)
print(f"leaves {tot}: duplicates within a pass {within / tot:.3%}")
for k, v in hits.items():
    print(f"  unique leaves already evaluated in the previous {k:5d} passes: {v / tot:.3%}")
