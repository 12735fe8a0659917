"""How fast is the CPU arm of bench.py (the oracle *port*) relative to the UNMODIFIED reference?

Runs only where /root/reference exists (the build container).  Times, at BASELINE config 1/2 settings
(11x11, 500 sims/move, upper 642, training-mode search from the empty board):

  single  : one process, reference `genData.player.Player(config, training=True, pv_fn=net.eval)` vs the port
            `oracle.mcts.OraclePlayer` with the same torch-CPU fp32 net (`oracle.net.OracleNet`, 1 thread)
  pipe5   : the reference's process topology (main.py:50-55): the reference's own `NetworkAPI` thread in the
            parent + 5 processes each running the reference `Player(pipe=...)`, vs bench.cpu_pipe_topology (port)

and writes profiles/r02_cpu_calibration.json with the port/reference ratios, so the `cpu_baseline` of bench.py
(kind "port") can be read as a statement about the reference itself.  TensorFlow is absent, so the net is the
oracle restatement in both arms; everything else on the reference side is the reference's own code.

    python tools/cpu_calibration.py [moves_single] [moves_per_worker]
"""
import contextlib
import json
import multiprocessing as mp
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"


class _Model:
    """What NetworkAPI needs from `agent_model` (networkAPI.py:67-68): `.graph.as_default()` and `.eval`."""

    class _G:
        def as_default(self):
            return contextlib.nullcontext()

    def __init__(self, S, threads):
        import torch
        from oracle import net as onet
        torch.set_num_threads(threads)
        self.graph = self._G()
        self.net = onet.OracleNet(S, onet.glorot_weights(S, 0))

    def eval(self, x):
        return self.net.eval(x)


def _ref_modules():
    sys.path.insert(0, REF)
    import config
    import utils
    from genData.player import Player
    return config, utils, Player


def ref_single(moves, sims, upper, seed=0):
    config, utils, Player = _ref_modules()
    config.board_size, config.simulation_per_step, config.upper_simulation_per_step = 11, sims, upper
    model = _Model(11, 1)
    np.random.seed(seed); random.seed(seed)
    pl = Player(config, training=True, pv_fn=model.eval)
    state, last = pl.get_init_state(), None
    t0 = time.perf_counter()
    for _ in range(moves):
        _, action = pl.get_action(state, last_action=last)
        board = utils.step(utils.state_to_board(state, 11), action)
        state, last = utils.board_to_state(board), action
    return moves / (time.perf_counter() - t0)


def port_single(moves, sims, upper, seed=0):
    from oracle import mcts, rules
    model = _Model(11, 1)
    cfg = mcts.SearchConfig(board_size=11, simulation_per_step=sims, upper_simulation_per_step=upper)
    pl = mcts.OraclePlayer(cfg, training=True, pv_fn=model.eval, rng=np.random.default_rng(seed))
    board, last = np.zeros((11, 11), np.int8), None
    t0 = time.perf_counter()
    for _ in range(moves):
        _, action = pl.get_action(board, last)
        board, last = rules.play(board, action), action
    return moves / (time.perf_counter() - t0)


def _ref_worker(pipe, q, moves, sims, upper, seed):
    """main.gen_data (main.py:82-94) cut to `moves` moves: the reference Player in pipe mode."""
    config, utils, Player = _ref_modules()
    config.board_size, config.simulation_per_step, config.upper_simulation_per_step = 11, sims, upper
    np.random.seed(seed); random.seed(seed)                 # fork copies numpy's global state
    pl = Player(config, training=True, pipe=pipe)
    state, last = pl.get_init_state(), None
    q.put("ready")
    for _ in range(moves):
        _, action = pl.get_action(state, last_action=last)
        board = utils.step(utils.state_to_board(state, 11), action)
        state, last = utils.board_to_state(board), action
    q.put(moves)


def ref_pipe5(moves_each, sims, upper, workers=5, net_threads=3):
    config, utils, Player = _ref_modules()
    from genData.networkAPI import NetworkAPI
    config.board_size = 11
    api = NetworkAPI(config, _Model(11, net_threads))
    ctx = mp.get_context("fork")                            # what the reference gets on Linux
    q = ctx.Queue()
    pipes = [api.get_pipe() for _ in range(workers)]
    procs = [ctx.Process(target=_ref_worker, args=(pipes[i], q, moves_each, sims, upper, 100 + i), daemon=True)
             for i in range(workers)]
    for p in procs:
        p.start()
    api.start(True)
    for _ in range(workers):
        assert q.get() == "ready"
    t0 = time.perf_counter()
    done = sum(q.get() for _ in range(workers))
    wall = time.perf_counter() - t0
    api.done = True
    for p in procs:
        p.join(timeout=5)
    return done / wall


def main():
    moves_single = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    moves_each = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    sims, upper = 500, 642
    import bench
    host = os.cpu_count()
    r1 = ref_single(moves_single, sims, upper)
    p1 = port_single(moves_single, sims, upper)
    net_threads = max(1, min(host - 5, 8))
    r5 = ref_pipe5(moves_each, sims, upper, net_threads=net_threads)
    p5, _, _ = bench.cpu_pipe_topology(11, sims, upper, 5, moves_each, net_threads=net_threads)
    out = {
        "where": f"build container, {host} host cores, numpy {np.__version__}; net = oracle.net.OracleNet (torch CPU fp32) in both arms "
                 "(TensorFlow 1.x, which the reference's network.py needs, is absent)",
        "settings": {"board": 11, "sims": sims, "upper": upper, "training": True, "from": "empty board"},
        "reference_single_moves_per_s": r1, "port_single_moves_per_s": p1, "port_over_reference_single": p1 / r1,
        "reference_pipe5_moves_per_s": r5, "port_pipe5_moves_per_s": p5, "port_over_reference_pipe5": p5 / r5,
        "moves_single": moves_single, "moves_per_worker": moves_each, "net_threads_pipe": net_threads,
        "note": "the port (array MCTS, bitboard-free numpy rules) is faster than the reference's dict/string Player; "
                "divide bench.py's cpu_baseline by port_over_reference_* to estimate the unmodified reference on the same cores",
    }
    path = os.path.join(ROOT, "profiles", "r02_cpu_calibration.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
