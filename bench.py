#!/usr/bin/env python
"""bench.py -- self-play moves/sec at 11x11, 500 sims/move (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this framework (B200)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port)

A *step* is `sims` lock-step passes (network forward over all N pending leaves + one tree
kernel pass) over the N concurrent games of this rank -- on average one move per game -- followed by
the harvest of the finished games, the all-gather of their records over NCCL (N > 1, side stream,
overlapped with the next step's passes) and their push into the device-resident RandomStack (the sink of
main.py:60-64) on every rank.
`value` is whole-job moves/s with everything resident in HBM (games are played, recorded
and restarted on the device); `e2e` is the same metric through the host-buffer API over a steady-state
mix of positions: BatchedPlayer.start_stream / poll / submit -- the searches that have ended are read back
to the host (policies, actions, next boards), the host sends the next roots, the other players keep
searching -- for max(--steps, 8) moves per player; `e2e.lockstep_call` is the one-move-of-every-player
call (BatchedPlayer.get_actions), which returns when the slowest search is done.  The other BASELINE
configurations (15x15 / 800 sims, the arena, a single game) are reported under `configs`.
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_LEAF = {11: 118_727_264, 15: 221_522_528}     # 2*MAC, biases/activations excluded
CONV_MAC_PER_CELL = 485_376          # the ten 3x3(+1x1) block convs, MACs per board cell (58,730,496 @ 11x11)
METRIC = "self-play moves/sec at 11x11, 500 sims/move; NN leaf-evals/sec"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--games", type=int, default=4096, help="concurrent games per GPU")
    ap.add_argument("--board", type=int, default=11)
    ap.add_argument("--sims", type=int, default=500)
    ap.add_argument("--upper", type=int, default=None)
    ap.add_argument("--net-mode", default=os.environ.get("A5_NET_MODE", "auto"), choices=["auto", "fp32", "tc"])
    ap.add_argument("--e2e-steps", type=int, default=None,
                    help="moves per player timed in the end-to-end leg (default: max(--steps, 8); 0 disables it)")
    ap.add_argument("--stream-passes", type=int, default=4, help="search passes queued per poll of the continuous e2e leg")
    ap.add_argument("--lockstep-calls", type=int, default=3, help="timed BatchedPlayer.get_actions calls (secondary e2e figure)")
    ap.add_argument("--preroll-moves", type=int, default=48,
                    help="untimed moves at --preroll-sims before the warm-up, so games are spread over all plies")
    ap.add_argument("--preroll-sims", type=int, default=40)
    ap.add_argument("--buffer", type=int, default=12000, help="RandomStack length in plies (config.buffer_size)")
    ap.add_argument("--accounting-passes", type=int, default=300)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg4 / cfg5 / cfg1 legs")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--eval-cache", type=int, default=int(os.environ.get("A5_EVAL_CACHE", "1")),
                    help="1: cross-game evaluation cache + compact leaf batch in the lock-step self-play runs (a5_evalcache_*); "
                         "it is shared with the continuous end-to-end leg; 2: also in the lock-step get_actions calls (there every "
                         "search of a call starts together, demand exceeds the compact batch in some calls and the stragglers "
                         "cost what the smaller forward gains: measured equal on average, profiles/r02_evalcache.txt); 0: off")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tensor": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm": 6650.0, "tensor": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].startswith("Active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(self.rows)}


# --------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference loop on the host cores.  Two topologies:
#   * "port": one independent process per core (numpy MCTS + torch-CPU fp32 net in-process, pv_fn mode)
#   * "pipe": the reference's own topology (main.py:50-55, networkAPI.py:43-78): a parent thread batches the
#     leaf requests of `workers` child processes over multiprocessing Pipes and evaluates them with one net
# profiles/r02_cpu_calibration.json relates both to the UNMODIFIED reference Player timed in the build container.
# --------------------------------------------------------------------------------------
def _cpu_worker(args):
    """Plays `moves` self-play moves (sims simulations each, training mode) with the oracle
    Player + oracle net; returns (moves, leaf_evals, seconds)."""
    S, sims, upper, moves, seed = args
    import numpy as np
    import torch
    torch.set_num_threads(1)
    from oracle import mcts, net, rules
    cfg = mcts.SearchConfig(board_size=S, simulation_per_step=sims, upper_simulation_per_step=upper)
    model = net.OracleNet(S, net.glorot_weights(S, 0))
    pl = mcts.OraclePlayer(cfg, training=True, pv_fn=model.eval, rng=np.random.default_rng(seed))
    board, last = np.zeros((S, S), np.int8), None
    t0 = time.perf_counter()
    done = 0
    for _ in range(moves):
        _, action = pl.get_action(board, last)
        board, last = rules.play(board, action), action
        done += 1
        if rules.terminal(board)[0]:
            pl.reset()
            board, last = np.zeros((S, S), np.int8), None
    return done, pl.stat_leaf_evals, time.perf_counter() - t0


def cpu_moves_per_sec(S, sims, upper, workers, moves_each, seed0=0):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(workers) as pool:
        res = pool.map(_cpu_worker, [(S, sims, upper, moves_each, seed0 + i) for i in range(workers)])
    wall = time.perf_counter() - t0
    busy = max(r[2] for r in res)
    moves, evals = sum(r[0] for r in res), sum(r[1] for r in res)
    return moves / busy, evals / busy, wall


def _pipe_worker(pipe, q, S, sims, upper, moves, seed):
    """main.gen_data's role (main.py:82-94) with the oracle Player: every leaf goes to the parent over the
    Pipe (player.py:194-197: send [x], spin on poll, recv()[0])."""
    import numpy as np
    from oracle import mcts, rules

    def pv(x):
        pipe.send([x[0]])
        while not pipe.poll():
            pass
        p, v = pipe.recv()[0]
        return p[None], np.float32([v])

    cfg = mcts.SearchConfig(board_size=S, simulation_per_step=sims, upper_simulation_per_step=upper)
    pl = mcts.OraclePlayer(cfg, training=True, pv_fn=pv, rng=np.random.default_rng(seed))
    board, last = np.zeros((S, S), np.int8), None
    q.put("ready")
    t0 = time.perf_counter()
    for _ in range(moves):
        _, action = pl.get_action(board, last)
        board, last = rules.play(board, action), action
    q.put((moves, pl.stat_leaf_evals, time.perf_counter() - t0))


def cpu_pipe_topology(S, sims, upper, workers, moves_each, net_threads, seed0=0):
    """The reference's process topology: `workers` processes + one inference thread in the parent
    (networkAPI.py:43-78: wait 1 ms, drain ready pipes, one batched eval, reply per pipe)."""
    import multiprocessing as mp
    from multiprocessing import connection
    import numpy as np
    import torch
    from oracle import net as onet
    torch.set_num_threads(max(1, net_threads))
    model = onet.OracleNet(S, onet.glorot_weights(S, 0))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    pipes, procs = [], []
    for i in range(workers):
        me, you = ctx.Pipe()
        pipes.append(me)
        p = ctx.Process(target=_pipe_worker, args=(you, q, S, sims, upper, moves_each, seed0 + i), daemon=True)
        p.start()
        procs.append(p)
    done = [False]

    def serve():
        while not done[0]:
            ready = connection.wait(pipes, timeout=0.001)
            if not ready:
                continue
            data, owners = [], []
            for pp in ready:
                try:
                    while pp.poll():
                        msg = pp.recv()
                        data.extend(msg)
                        owners.append((pp, len(msg)))
                except (EOFError, OSError):
                    if pp in pipes:
                        pipes.remove(pp)
            if not data:
                continue
            prob, value = model.eval(np.asarray(data, np.float32))
            k = 0
            for pp, n in owners:
                pp.send([(prob[k + j], float(value[k + j])) for j in range(n)])
                k += n

    th = threading.Thread(target=serve, daemon=True)
    th.start()
    for _ in range(workers):
        assert q.get() == "ready"
    t0 = time.perf_counter()
    res = [q.get() for _ in range(workers)]
    wall = time.perf_counter() - t0
    done[0] = True
    for p in procs:
        p.join(timeout=5)
    return sum(r[0] for r in res) / wall, sum(r[1] for r in res) / wall, wall


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S, sims = a.board, a.sims
    upper = a.upper or sims + 142
    cores = os.cpu_count() or 1
    workers = max(1, min(cores - 1, 64))
    for _ in range(a.warmup and 1):                       # one warm-up round is enough for a CPU loop
        cpu_moves_per_sec(S, sims, upper, workers, 1, seed0=1000)
    t0 = time.perf_counter()
    tot_m = tot_e = 0.0
    for k in range(a.steps):
        m, e, _ = cpu_moves_per_sec(S, sims, upper, workers, 1, seed0=100 * k)
        tot_m += m
        tot_e += e
    wall = time.perf_counter() - t0
    v = tot_m / a.steps
    sample = (f"{workers} processes x 1 move of {sims} sims per step from the empty board, training mode, oracle port + "
              f"torch-CPU fp32 net (1 thread each); host has {cores} cores")
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "moves/s", "leaf_evals_per_s": tot_e / a.steps,
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000 * wall / max(1, a.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a, max(1, a.gpus)), "board": S, "sims": sims, "upper_sims": upper},
        "cpu_baseline": {"value": v, "unit": "moves/s", "cores": workers, "host_cores": cores, "kind": "port", "sample": sample,
                         "calibration": calibration_note()},
        "e2e": {"value": v, "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def calibration_note():
    p = os.path.join(ROOT, "profiles", "r02_cpu_calibration.json")
    try:
        d = json.load(open(p))
        return {k: d[k] for k in ("port_over_reference_single", "port_over_reference_pipe5", "where") if k in d}
    except Exception:
        return None


def workload_name(a, world):
    return (f"{a.games * world} concurrent {a.board}x{a.board} games ({a.games}/GPU), {a.sims} sims/move, "
            f"random-init net (glorot-uniform, seed 0), training-mode self-play")


# --------------------------------------------------------------------------------------
def conv_roofline(kt, S, N, pk, traffic):
    """N = boards per network forward (the compact batch when the evaluation cache is on).  Dominant kernel = k_tc_conv2 (eight launches per pass: block3/block4 conv1 run as one layer, and so do
    block3/block4 conv2; 98.9 % of the algorithmic FLOPs).  achieved = algorithmic conv FLOPs of one pass /
    in-situ time of those launches (predecessor's end -> own end inside the CUDA-graph replay)."""
    C = S * S
    names = [f"k_tc_conv2[{i}]" for i in range(8)] + ["k_tc_mega"]
    conv_us = sum(kt[n][0] for n in names if n in kt)
    mega = "k_tc_mega" in kt
    conv_flop = 2.0 * CONV_MAC_PER_CELL * C * N
    ach = conv_flop / (conv_us * 1e-6) / 1e12
    net_us = sum(v[0] for k, v in kt.items() if k not in ("k_step", "(fold)", "k_ec_lookup", "k_ec_commit"))
    flop = FLOP_PER_LEAF.get(S, FLOP_PER_LEAF[11] * C / 121) * N
    tr = traffic.get("conv_dram_bytes_per_pass") * N / float(traffic.get("boards", 4096)) * (C / 121.0) if traffic else None
    return {"bound": "tensor", "kernel": ("k_tc_mega (the eight block-conv layers as one chunk-major tcgen05 cta_group::2 launch)" if mega else
                                          "k_tc_conv2 (tcgen05 cta_group::2 3x3 conv + residual, 8 launches per pass)"),
            "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": ach / pk["tensor"],
            "peak_source": pk["src"] + " bf16 sustained (MEASURED_PEAKS.json)",
            "timing": "in situ: %globaltimer stamps of every launch inside CUDA-graph replays of the running workload; "
                      "a kernel's time = its predecessor's end to its own end, so the kernels partition the pass",
            "traffic": tr,
            "traffic_note": (traffic.get("note") + f"; scaled to {N} boards") if traffic else "no ncu capture committed for this build",
            "algorithmic_flop_per_launch_group": conv_flop, "split_passes_issued": 3, "us_per_pass": conv_us,
            "us_layers": {n: round(kt[n][0], 2) for n in names if n in kt},
            "whole_net": {"kernel": "bitboards + conv1 + 8 block-conv launches + heads (11 launches)", "us_per_pass": net_us,
                          "achieved": flop / (net_us * 1e-6) / 1e12, "frac": flop / (net_us * 1e-6) / 1e12 / pk["tensor"]}}


def tree_roofline(kt, c0, c1, S, pk):
    """k_step against the HBM roofline with SURVEY 8(d)'s algorithmic bytes and the run's measured d, A, f_leaf."""
    C = S * S
    nsims = max(1.0, c1["sims"] - c0["sims"])
    dbar = (c1["selects"] - c0["selects"]) / nsims
    abar = (c1["legal_sum"] - c0["legal_sum"]) / max(1.0, c1["leaf_evals"] - c0["leaf_evals"])
    fleaf = (c1["leaf_evals"] - c0["leaf_evals"]) / nsims
    bytes_per_sim = dbar * (12 * abar + 16 + C) + dbar * 16 + dbar * 20 + fleaf * (3 * C + 4 * C + 4 + 12 * abar + C + 16)
    sims_per_pass = nsims / max(1.0, c1["passes"] - c0["passes"])
    us = kt["k_step"][0]
    gbs = bytes_per_sim * sims_per_pass / (us * 1e-6) / 1e9
    return {"bound": "hbm", "kernel": "k_step (tree pass)", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
            "frac": gbs / pk["hbm"], "us_per_launch": us, "bytes_per_sim": bytes_per_sim, "sims_per_launch": sims_per_pass,
            "d_bar": dbar, "a_bar": abar, "f_leaf": fleaf}


def measured_traffic():
    """DRAM bytes per pass of the block convs from the committed ncu capture (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def desync_budgets(engine, sims, seed):
    """Spread the phases of the searches uniformly: every game's *current* move gets a remaining budget drawn from
    [1, sims] (warm-up only).  Without it all games switch to the full budget in the same pass after the preroll (or
    start their first search together in the end-to-end leg), finish their moves in waves sims passes apart, and
    the number of moves inside the timed window depends on where its edges fall between two waves (+-2 %)."""
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(int(seed))
    sl = engine.sims_left()
    sl.copy_(torch.minimum(sl, torch.randint(1, sims + 1, sl.shape, device=sl.device, dtype=torch.int32, generator=g)))


# --------------------------------------------------------------------------------------
def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from alphafive_b200 import _lib
    from alphafive_b200.net import DeviceNet, glorot_init
    from alphafive_b200.replay import gather_records
    from alphafive_b200.replay_stack import RandomStack
    from alphafive_b200.selfplay import BatchedPlayer, SelfPlay

    S, N, sims = a.board, a.games, a.sims
    upper = a.upper or sims + 142                           # config.py:4-5 keeps upper = sims + 100..142
    lib = _lib.load()
    mode = {"fp32": _lib.NET_FP32, "tc": _lib.NET_TC}.get(a.net_mode)
    if mode is None:
        mode = _lib.NET_TC if getattr(lib, "a5_net_tc_available", lambda: 0)() else _lib.NET_FP32
    weights = glorot_init(S, 0)
    net = DeviceNet(S, N, weights, mode=mode)
    sp = SelfPlay(None, n_games=N, net=net, training=True, seed=0, game_id_base=rank * N, use_graph=not a.no_graph,
                  eval_cache=(dict(cap=int(os.environ["A5_EC_CAP"])) if os.environ.get("A5_EC_CAP") else True) if (a.eval_cache and mode == _lib.NET_TC) else False,
                  board_size=S, simulation_per_step=sims, upper_simulation_per_step=upper)
    stride = sp.engine.record_stride
    stack = RandomStack(S, length=a.buffer)                  # the sink: every rank keeps the whole replay buffer
    bufs = [torch.empty_like(sp.record_buf) for _ in range(2)]
    side = torch.cuda.Stream()
    sink = {"pending": None, "k": 0, "records": 0, "games_pushed": 0, "accepted": 0, "gather_bytes": 0, "gather_ms": 0.0,
            "events": []}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def drain_sink():
        """Gather + push what the previous step harvested; runs on the side stream while the main stream
        works through the passes already enqueued."""
        buf = sink["pending"]
        if buf is None:
            return
        sink["pending"] = None
        with torch.cuda.stream(side):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            allrec, counts = gather_records(buf)             # count-sized: counts first, then max-count padded payload
            e1.record()
            flags = stack.push_records(allrec)
            sink["events"].append((e0, e1))
            sink["gather_bytes"] += int(counts.max()) * stride * world if world > 1 else 0
            sink["records"] += int(allrec.shape[0])
            sink["games_pushed"] += len(flags)
            sink["accepted"] += sum(flags)

    def one_step():
        sp.run_passes(sims)                                   # enqueued; the host is free while they run
        drain_sink()
        torch.cuda.current_stream().wait_stream(side)         # (after the queued passes) the buffer about to be reused is free
        buf, games = sp.harvest(bufs[sink["k"] % 2])          # synchronises the main stream
        sink["k"] += 1
        side.wait_stream(torch.cuda.current_stream())         # the gather must see the harvested records
        sink["pending"] = buf
        return buf.shape[0], games

    sp.start()
    if a.preroll_moves > 0:
        # Desynchronise the games: a short-budget prefix plays ~preroll_moves plies per game (games
        # finish and restart on the way), so the warm-up and the timed steps see the steady-state mix
        # of openings, middle games, terminal positions and record emission -- not 4096 empty boards.
        sp.set_budget(a.preroll_sims, a.preroll_sims + 10)
        sp.run_passes(a.preroll_moves * a.preroll_sims)
        sp.harvest()
        sp.set_budget(sims, upper)
        sp.run_passes(a.preroll_sims + 12)                    # every game has started a move with the full budget
        desync_budgets(sp.engine, sims, 1234 + rank)
    for _ in range(a.warmup):
        one_step()
    drain_sink()
    side.synchronize()
    for key in ("records", "games_pushed", "accepted", "gather_bytes"):
        sink[key] = 0
    sink["events"] = []
    c0 = sp.counters()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    recs = 0
    for _ in range(a.steps):
        r, _ = one_step()
        recs += r
    drain_sink()                                              # the last step's records reach the buffer inside the timed region
    torch.cuda.current_stream().wait_stream(side)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    c1 = sp.counters()
    clocks = sampler.stop() if rank == 0 else None
    gather_ms = sum(x.elapsed_time(y) for x, y in sink["events"])

    # ---- in-situ kernel accounting: the same workload, CUDA-graph replays with %globaltimer stamps ------
    kt = sp.kernel_accounting(a.accounting_passes) if mode == _lib.NET_TC and not a.no_graph else None
    c2 = sp.counters()

    t = torch.tensor([ms, float(c1["moves"] - c0["moves"]), float(c1["leaf_evals"] - c0["leaf_evals"]),
                      float(c1["sims"] - c0["sims"]), float(c1["games"] - c0["games"])], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, moves, evals, nsims, games = tmax[0].item(), tsum[1].item(), tsum[2].item(), tsum[3].item(), tsum[4].item()
    else:
        ms, moves, evals, nsims, games = [x.item() for x in t]
    value = moves / (ms / 1000)
    passes_timed = c1["passes"] - c0["passes"]

    # ---- end to end through the host-buffer API (BatchedPlayer.get_actions), steady state ----------------
    # Start from where the lock-step games stand (a stationary mix of plies), let a warm-up call build the
    # tables, then time consecutive moves: tree reuse, budget cuts, terminal positions and restarts included.
    e2e_steps = a.e2e_steps if a.e2e_steps is not None else max(a.steps, 8)
    d_boards, d_last = sp.engine.roots()
    tau0 = sp.engine.tau().clone()                           # every game's temperature (Player.tau): the end-to-end legs resume these games
    boards, last = d_boards.cpu().numpy(), d_last.cpu().numpy()
    occupancy = stack.count / stack.length
    cache_info = None
    if sp.cache is not None:
        st = sp.cache.stats()
        cache_info = {"boards_per_forward": sp.cache.cap, "of_games": N, "lookups": st["lookups"],
                      "hit_rate": st["hits"] / max(1, st["lookups"]), "deferred_rate": st["deferred"] / max(1, st["lookups"]),
                      "stored": st["stored"], "_hits": st["hits"], "_deferred": st["deferred"],
                      "note": "since the start of the run (preroll and warm-up included); leaves served from earlier network "
                              "results of any game, the rest evaluated as a compact batch (a5_evalcache_*); per-game results "
                              "are bit-identical to the uncached run (tests/test_gpu_mcts.py)"}
    shared_cache = sp.cache                                  # warm: the end-to-end leg below searches with the same weights
    launches_per_pass = (12 if mode == _lib.NET_TC else 16) + (3 if cache_info is not None else 0)
    sp.engine.close()
    del sp, stack, bufs
    torch.cuda.empty_cache()
    e2e_val, e2e_calls, bp_bytes, e2e_cache, e2e_stream = None, [], (0, 0), None, None
    if e2e_steps > 0:
        # (1) continuous batching (BatchedPlayer.start_stream / poll / submit): every search that has ended is read
        # back, stepped on the host side of the API and re-rooted while the others keep searching -- the headline e2e.
        bp = BatchedPlayer(None, n_players=N, net=net, training=True, seed=1, game_id_base=rank * N,
                           eval_cache=shared_cache if shared_cache is not None else False,
                           board_size=S, simulation_per_step=sims, upper_simulation_per_step=upper)
        bp.start_stream(boards, last, np.ones(N, np.uint8), passes=a.stream_passes)
        bp.engine.tau().copy_(tau0)                              # Player.tau of the resumed games (a fresh Player starts at init_temp)
        desync_budgets(bp.engine, sims, 4321 + rank)             # (first search of every player only)

        def stream_until(target):
            got = 0
            while got < target:
                games, pol, act, nxt, codes = bp.poll()
                over = codes != 0
                bp.submit(games, np.where(over[:, None, None], 0, nxt).astype(np.int8), np.where(over, -1, act).astype(np.int32),
                          over.astype(np.uint8))                 # a finished game restarts: Player.reset() (player.py:73)
                got += len(games)
            return got

        stream_until(2 * N)                                      # warm-up: tables built, budgets desynchronised
        st0 = bp.cache.stats() if bp.cache is not None else None
        b0 = bp.stream_stats()
        ec0 = bp.engine.counters()
        barrier()
        t0 = time.perf_counter()
        got = stream_until(e2e_steps * N)
        t_stream = time.perf_counter() - t0
        barrier()
        tt = torch.tensor([t_stream, float(got)], dtype=torch.float64, device="cuda")
        if world > 1:
            tmx = tt.clone(); dist.all_reduce(tmx, op=dist.ReduceOp.MAX)
            tsm = tt.clone(); dist.all_reduce(tsm, op=dist.ReduceOp.SUM)
            t_stream, got_all = tmx[0].item(), tsm[1].item()
        else:
            got_all = float(got)
        e2e_val = got_all / t_stream
        b1 = bp.stream_stats()
        polls = b1["polls"] - b0["polls"]
        ec1 = bp.engine.counters()
        steps_equiv = max(1e-9, got / N)                         # one "step" = N moves, as in the lock-step call
        e2e_stream = {"moves": got_all, "seconds": t_stream, "polls": polls, "passes_per_poll": a.stream_passes,
                      "collect_cap": b1["collect_cap"], "moves_per_poll": got / max(1, polls),
                      "sims_per_move": (ec1["sims"] - ec0["sims"]) / max(1, ec1["moves"] - ec0["moves"]),
                      "passes_per_move_per_player": N * polls * a.stream_passes / max(1, got),
                      "h2d_bytes_per_step": (b1["h2d_bytes"] - b0["h2d_bytes"]) / steps_equiv,
                      "d2h_bytes_per_step": (b1["d2h_bytes"] - b0["d2h_bytes"]) / steps_equiv}
        if bp.cache is not None:
            st = bp.cache.stats()
            lk = max(1, st["lookups"] - st0["lookups"])
            e2e_cache = {"lookups": lk, "hit_rate": (st["hits"] - st0["hits"]) / lk, "deferred_rate": (st["deferred"] - st0["deferred"]) / lk}
        bp.engine.close()
        del bp
        torch.cuda.empty_cache()
        # (2) the lock-step call (BatchedPlayer.get_actions): one call = one move of every player, returns when the
        # slowest search is done
        bp = BatchedPlayer(None, n_players=N, net=net, training=True, seed=1, game_id_base=rank * N,
                           eval_cache=shared_cache if (shared_cache is not None and a.eval_cache >= 2) else False,
                           board_size=S, simulation_per_step=sims, upper_simulation_per_step=upper)
        bp_bytes = (bp.h2d_bytes, bp.d2h_bytes)
        bp.engine.tau().copy_(tau0)
        clear = np.zeros(N, np.uint8)                        # fresh engine: the tables are empty, tau is the resumed games'
        n_calls = max(2, min(e2e_steps, a.lockstep_calls))
        for k in range(2 + n_calls):
            timed = k >= 2
            barrier()
            t0 = time.perf_counter()
            pol, act, nxt, codes = bp.get_actions(boards, last, None, clear, advance=True)
            over = codes != 0
            boards = np.where(over[:, None, None], 0, nxt).astype(np.int8)
            last = np.where(over, -1, act).astype(np.int32)
            clear = over.astype(np.uint8)                    # a finished game restarts: Player.reset() (player.py:73)
            barrier()
            if timed:
                e2e_calls.append(time.perf_counter() - t0)
        e2e_t = torch.tensor([sum(e2e_calls)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e_lock = N * world * len(e2e_calls) / e2e_t.item()
        bp.engine.close()
        del bp
        torch.cuda.empty_cache()
    if shared_cache is not None:
        shared_cache.close()

    out = None
    if rank == 0:
        pk = peaks()
        out = {
            "metric": METRIC, "value": value, "unit": "moves/s", "leaf_evals_per_s": evals / (ms / 1000),
            "leaf_evals_note": ("leaves expanded with a network result per second; with the evaluation cache a pass sends "
                                "`eval_cache.boards_per_forward` boards through the network and serves the other leaves from "
                                "earlier results" if cache_info else "leaves expanded = boards through the network"),
            "nn_boards_per_s": (cache_info["boards_per_forward"] if cache_info else N) * world * passes_timed / (ms / 1000),
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if mode == _lib.NET_FP32 else "f16x2-split (fp32 accumulate)", "data": "synthetic",
            "config": {"workload": workload_name(a, world), "board": S, "sims": sims, "upper_sims": upper,
                       "games_per_gpu": N, "net_mode": "fp32" if mode == _lib.NET_FP32 else "tc",
                       "l2": "per-pass working set (activations + node arenas) exceeds the 126 MB L2",
                       "cuda_graph": not a.no_graph, "step": f"{sims} lock-step passes + harvest + record all-gather + RandomStack push",
                       "schedule": "single stream, CUDA-graph replay of one pass; record gather + replay-buffer push on a side stream",
                       "preroll": f"{a.preroll_moves} untimed moves at {a.preroll_sims} sims to spread games over all plies, then the "
                                  "remaining budgets of the moves in progress are drawn uniformly so that the searches end at "
                                  "uniformly spread passes (no waves of moves across the timed window's edges)"},
            "e2e": ({"value": e2e_val, "unit": "moves/s", "h2d_bytes_per_step": e2e_stream["h2d_bytes_per_step"],
                     "d2h_bytes_per_step": e2e_stream["d2h_bytes_per_step"],
                     "api": "BatchedPlayer.start_stream / poll / submit (continuous batching over a5_engine_collect_moves / "
                            "a5_engine_submit_roots): the searches that have ended are read back to the host (policy, action, "
                            "next position, terminal code), the host restarts finished games and sends the next roots, the other "
                            "players keep searching; positions and temperatures from the lock-step run (tree reuse, restarts); one "
                            "step = as many moves as there are players",
                     "stream": e2e_stream,
                     "lockstep_call": {"value": e2e_lock, "unit": "moves/s", "h2d_bytes_per_step": bp_bytes[0], "d2h_bytes_per_step": bp_bytes[1],
                                       "api": "BatchedPlayer.get_actions(host boards) + device step/terminal: one move of every player "
                                              "per call, returns when the slowest search is done",
                                       "calls": len(e2e_calls), "seconds": sum(e2e_calls),
                                       "moves_per_s_min": N * world / max(e2e_calls), "moves_per_s_max": N * world / min(e2e_calls)}}
                    if e2e_stream is not None else {"value": None, "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}),
            "gpu_launches": int(passes_timed * launches_per_pass),
            "ms_per_pass": ms / max(1.0, passes_timed),
            "moves": moves, "sims_run": nsims, "games_finished": games,
            "replay": {"records_per_step": sink["records"] / max(1, a.steps), "local_records_per_step": recs / max(1, a.steps),
                       "games_pushed_per_step": sink["games_pushed"] / max(1, a.steps),
                       "games_accepted_per_step": sink["accepted"] / max(1, a.steps),
                       "allgather_bytes_per_step": sink["gather_bytes"] / max(1, a.steps),
                       "collective_us_per_step": 1000.0 * gather_ms / max(1, a.steps),
                       "ring_occupancy": occupancy, "buffer_plies": a.buffer,
                       "sink": "RandomStack.push_records on every rank (utils.py:64-116 decisions, device ring)"},
            "clocks": clocks,
        }
        if cache_info is not None:
            out["eval_cache"] = {k: v for k, v in cache_info.items() if not k.startswith("_")}
            if e2e_cache is not None:
                out["eval_cache"]["e2e_leg"] = e2e_cache
        if kt is not None:
            out["kernels_us_per_pass"] = {k: [round(v[0], 2), round(v[1], 2)] for k, v in kt.items()}
            out["kernels_note"] = ("[predecessor end -> own end, own first start -> own end] per kernel, mean over "
                                   f"{a.accounting_passes} CUDA-graph replays right after the timed region; the first values sum to "
                                   "the accounted pass")
            n_fwd = cache_info["boards_per_forward"] if cache_info else N
            out["roofline"] = conv_roofline(kt, S, n_fwd, pk, measured_traffic())
            out["roofline"]["boards_per_forward"] = n_fwd
            out["roofline_tree"] = tree_roofline(kt, c1, c2, S, pk)
            # the stamps cost ~1 % (atomics, one extra CTA barrier per kernel) and the fold kernel is instrumentation only
            out["ms_per_pass_accounted"] = sum(v[0] for k, v in kt.items() if k != "(fold)") / 1000.0
            out["accounting_note"] = ("accounted = sum of the kernels of the instrumented replays without the stamp-folding "
                                      "kernel; instrumented replays run ~1 % slower than the timed ones")
        else:
            flop = FLOP_PER_LEAF.get(S, FLOP_PER_LEAF[11] * S * S / 121)
            tf = flop * evals / (ms / 1000) / 1e12
            out["roofline"] = {"bound": "tensor", "kernel": "policy/value net forward (whole pass, no per-kernel accounting)",
                               "achieved": tf, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": tf / pk["tensor"], "traffic": None}

    # ---- the other BASELINE configurations ------------------------------------------------------------
    if not a.no_configs:
        cfgs = run_configs(a, rank, world, net, mode)
        if rank == 0:
            out["configs"] = cfgs
    if rank == 0:
        if not a.no_cpu_baseline and world == 1:      # the CPU arm is reported at N = 1 only
            host = os.cpu_count() or 1
            cores = max(1, min(host - 1, 64))
            v, ev, wall = cpu_moves_per_sec(S, sims, upper, cores, 2)
            v5, ev5, wall5 = cpu_pipe_topology(S, sims, upper, 5, 2, net_threads=max(1, min(host - 5, 8)))
            out["cpu_baseline"] = {"value": v, "unit": "moves/s", "cores": cores, "host_cores": host, "kind": "port",
                                   "leaf_evals_per_s": ev,
                                   "five_workers": {"value": v5, "leaf_evals_per_s": ev5, "workers": 5,
                                                    "topology": "5 processes + one batching inference thread in the parent over "
                                                                "multiprocessing Pipes (main.py:50-55, networkAPI.py:43-78)",
                                                    "seconds": wall5},
                                   "calibration": calibration_note(),
                                   "sample": f"{cores} processes x 2 moves of {sims} sims from the empty board, oracle port "
                                             f"(numpy MCTS + torch-CPU fp32 net, 1 thread each), {wall:.1f}s wall"}
        emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_configs(a, rank, world, net11, mode):
    """BASELINE configs 4 (15x15, 800 sims), 5 (arena, 1024 paired games at 400 sims, sharded over the ranks) and
    1 (one game, B = 1 latency of Player.get_action).  Short legs: a few steps each."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from alphafive_b200 import _lib
    from alphafive_b200.net import DeviceNet, glorot_init
    from alphafive_b200.selfplay import SelfPlay
    pk = peaks()
    out = {}

    def reduce(vals, op):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=op)
        return t.tolist()

    # ---- cfg 4: 4096 concurrent 15x15 games, 800 sims/move -------------------------------------------
    try:
        S, N, sims, upper = 15, a.games, 800, 900
        net = DeviceNet(S, N, glorot_init(S, 0), mode=mode)
        sp = SelfPlay(None, n_games=N, net=net, training=True, seed=0, game_id_base=rank * N, board_size=S,
                      eval_cache=bool(a.eval_cache) and mode == _lib.NET_TC,
                      simulation_per_step=sims, upper_simulation_per_step=upper)
        sp.start()
        sp.set_budget(40, 50)
        sp.run_passes(32 * 40)
        sp.harvest()
        sp.set_budget(sims, upper)
        sp.run_passes(52)
        desync_budgets(sp.engine, sims, 99 + rank)
        sp.run_passes(sims)                                   # warm-up step
        sp.harvest()
        c0 = sp.counters()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        steps4 = 2
        for _ in range(steps4):
            sp.run_passes(sims)
            sp.harvest()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        c1 = sp.counters()
        kt = sp.kernel_accounting(100) if mode == _lib.NET_TC else None
        c2 = sp.counters()
        mx, = reduce([ms], dist.ReduceOp.MAX) if world > 1 else [ms]
        mv, ev = reduce([c1["moves"] - c0["moves"], c1["leaf_evals"] - c0["leaf_evals"]], dist.ReduceOp.SUM) if world > 1 else \
            [c1["moves"] - c0["moves"], c1["leaf_evals"] - c0["leaf_evals"]]
        out["cfg4"] = {"workload": f"{N * world} concurrent 15x15 games, 800 sims/move (upper 900: the reference keeps upper = sims + 100), "
                                   "random-init net, training-mode self-play", "moves_per_s": mv / (mx / 1000),
                       "leaf_evals_per_s": ev / (mx / 1000), "steps": steps4, "ms_per_step": mx / steps4,
                       "overflows": c1["overflows"], "max_nodes": c1["max_nodes"], "flop_per_leaf": FLOP_PER_LEAF[15]}
        if kt:
            n_fwd = sp.cache.cap if sp.cache is not None else N
            out["cfg4"]["roofline"] = conv_roofline(kt, S, n_fwd, pk, None)
            out["cfg4"]["roofline"]["boards_per_forward"] = n_fwd
            out["cfg4"]["roofline_tree"] = tree_roofline(kt, c1, c2, S, pk)
            out["cfg4"]["kernels_us_per_pass"] = {k: [round(v[0], 2), round(v[1], 2)] for k, v in kt.items()}
        sp.engine.close()
        net.close()
        del sp, net
        torch.cuda.empty_cache()
    except Exception as e:                                    # a side leg must not take the headline down
        out["cfg4"] = {"error": repr(e)}

    # ---- cfg 5: arena, 1024 paired games between two weight sets, 400 sims/move ----------------------
    try:
        import types
        from alphafive_b200 import config as cfgmod
        from alphafive_b200.drivers import Arena, count_wins
        from alphafive_b200.replay import allreduce_wins
        S, total, sims = 11, 1024, 400
        n_local = total // world
        z = np.load(os.path.join(ROOT, "tests", "golden", "ckpt6960.npz"))
        w_a = {k.replace("__", "/"): z[k] for k in z.files}
        net_a = DeviceNet(S, n_local, w_a, mode=mode)
        net_b = DeviceNet(S, n_local, glorot_init(S, 0), mode=mode)
        cfg = types.SimpleNamespace(**{k: getattr(cfgmod, k) for k in dir(cfgmod) if not k.startswith("_")})
        cfg.board_size, cfg.simulation_per_step, cfg.upper_simulation_per_step = S, sims, sims + 242   # choose_best_player.py:25: 400 of 642
        arena = Arena(cfg, net_a, net_b, n_local, seed=7, game_id_base=rank * n_local)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = arena.play()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        winners = res["winners"]
        w0, w1, dr = allreduce_wins(int((winners == 0).sum()), int((winners == 1).sum()), int((winners < 0).sum()), device="cuda")
        mx, = reduce([dt], dist.ReduceOp.MAX) if world > 1 else [dt]
        mv, ev = reduce([res["moves"], res["leaf_evals"]], dist.ReduceOp.SUM) if world > 1 else [res["moves"], res["leaf_evals"]]
        e0_, e1_, counted = count_wins(winners.tolist())      # the reference's sequential early-stop rule on this rank's games
        out["cfg5"] = {"workload": f"arena: {total} paired games to completion ({n_local}/GPU), {sims} sims/move, ckpt-6960 vs glorot "
                                   "seed 0, Player(training=False).get_action(random_a=True), separate table per player "
                                   "(choose_best_player.py:24-72)",
                       "wins_ckpt6960": w0, "wins_glorot": w1, "draws": dr, "moves_per_s": mv / mx, "leaf_evals_per_s": ev / mx,
                       "seconds": mx, "mean_plies": float(res["plies"].mean()),
                       "early_stop_rule_rank0": {"wins0": e0_, "wins1": e1_, "games_counted": counted},
                       "collective": "allreduce_wins (3 int64, NCCL)" if world > 1 else "none (1 GPU)"}
        arena.close()
        net_a.close(); net_b.close()
        del arena, net_a, net_b
        torch.cuda.empty_cache()
    except Exception as e:
        out["cfg5"] = {"error": repr(e)}

    # ---- cfg 1 on the GPU: one game, Player.get_action latency (B = 1) -------------------------------
    if rank == 0:
        try:
            import types
            from alphafive_b200 import config as cfgmod
            from alphafive_b200.genData.network import ResNet
            from alphafive_b200.genData.player import Player, board_to_state, state_to_board
            cfg = types.SimpleNamespace(**{k: getattr(cfgmod, k) for k in dir(cfgmod) if not k.startswith("_")})
            cfg.board_size, cfg.simulation_per_step, cfg.upper_simulation_per_step = 11, 500, 642
            rn = ResNet(11, max_batch=8, seed=0)
            pl = Player(cfg, training=False, pv_fn=rn.eval)
            state, lat = pl.get_init_state(), []
            from alphafive_b200 import utils as a5utils
            for ply in range(7):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                _, action = pl.get_action(state, last_action=None)          # self_play.py:95-97 resets last_action
                lat.append(time.perf_counter() - t0)
                board = a5utils.step(state_to_board(state, 11), action)
                state = board_to_state(board)
            pl.close()
            rn.close()
            out["cfg1_gpu"] = {"workload": "single 11x11 game, 500 sims/move, random-init net, Player(training=False).get_action "
                                           "with the on-device net (self_play.py:94-101)", "ms_per_move": 1000 * float(np.mean(lat[1:])),
                               "ms_first_move": 1000 * lat[0], "moves_per_s": 1.0 / float(np.mean(lat[1:])), "moves_timed": len(lat) - 1}
        except Exception as e:
            out["cfg1_gpu"] = {"error": repr(e)}
    return out


def emit(line: str):
    """The one JSON line goes to the process's real stdout (see main)."""
    os.write(_REAL_STDOUT, (line + "\n").encode())


_REAL_STDOUT = 1

if __name__ == "__main__":
    args = parse()
    # stdout carries exactly one JSON line: libraries that print there (NCCL writes its version banner to
    # stdout under NCCL_DEBUG=VERSION) are sent to stderr for the whole run
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
