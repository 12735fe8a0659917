#!/usr/bin/env python
"""bench.py -- self-play moves/sec at 11x11, 500 sims/move (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this framework (B200)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port)

A *step* is `sims` lock-step passes (network forward over all N pending leaves + one tree
kernel pass) over the N concurrent games of this rank: on average one move per game.
`value` is whole-job moves/s with everything resident in HBM (games are played, recorded
and restarted on the device); `e2e` is the same metric through the host-buffer API
(BatchedPlayer.get_actions: H2D boards, search, D2H policies/actions/next boards each
move).  One JSON line on stdout (rank 0).
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
METRIC = "self-play moves/sec at 11x11, 500 sims/move; NN leaf-evals/sec"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--games", type=int, default=4096, help="concurrent games per GPU")
    ap.add_argument("--board", type=int, default=11)
    ap.add_argument("--sims", type=int, default=500)
    ap.add_argument("--upper", type=int, default=None)
    ap.add_argument("--net-mode", default=os.environ.get("A5_NET_MODE", "auto"), choices=["auto", "fp32", "tc"])
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--preroll-moves", type=int, default=48,
                    help="untimed moves at --preroll-sims before the warm-up, so games are spread over all plies")
    ap.add_argument("--preroll-sims", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--pipeline", action="store_true",
                    help="two half batches pipelined on SM-partitioned streams (CUDA green contexts) instead of the "
                         "single-stream CUDA-graph replay; bit-identical results, measured no faster on a power-capped "
                         "B200 (DESIGN.md section 8)")
    ap.add_argument("--small-sms", type=int, default=16, help="SMs of the small partition of the pipeline")
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
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].startswith("Active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# --------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference loop, all host cores
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
    sample = f"{workers} processes x 1 move of {sims} sims per step from the empty board, training mode, oracle port + torch-CPU fp32 net (1 thread each)"
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "moves/s", "leaf_evals_per_s": tot_e / a.steps,
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000 * wall / max(1, a.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a, max(1, a.gpus)), "board": S, "sims": sims, "upper_sims": upper},
        "cpu_baseline": {"value": v, "unit": "moves/s", "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_name(a, world):
    return (f"{a.games * world} concurrent {a.board}x{a.board} games ({a.games}/GPU), {a.sims} sims/move, "
            f"random-init net (glorot-uniform, seed 0), training-mode self-play")


# --------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import ctypes as ct
    from alphafive_b200 import _lib
    from alphafive_b200._lib import check, ptr, stream_ptr
    from alphafive_b200.net import DeviceNet, glorot_init
    from alphafive_b200.selfplay import BatchedPlayer, SelfPlay

    S, N, sims = a.board, a.games, a.sims
    upper = a.upper or sims + 142                           # config.py:4-5 keeps upper = sims + 100..142
    lib = _lib.load()
    mode = {"fp32": _lib.NET_FP32, "tc": _lib.NET_TC}.get(a.net_mode)
    if mode is None:
        mode = _lib.NET_TC if getattr(lib, "a5_net_tc_available", lambda: 0)() else _lib.NET_FP32
    pipelined = mode == _lib.NET_TC and a.pipeline and N % 2 == 0
    part = None
    if pipelined:
        from alphafive_b200.pipeline import SmPartition
        from alphafive_b200.selfplay import PipelinedBatchedPlayer, PipelinedSelfPlay
        try:
            part = SmPartition.get(a.small_sms)
        except Exception as e:                                # driver without green contexts: single stream
            print(f"[bench] SM partition unavailable ({e}); single-stream schedule", file=sys.stderr)
            pipelined = False
    weights = glorot_init(S, 0)
    if pipelined:
        sp = PipelinedSelfPlay(None, n_games=N, weights=weights, training=True, seed=0, game_id_base=rank * N,
                               small_sms=a.small_sms, board_size=S, simulation_per_step=sims,
                               upper_simulation_per_step=upper)
        net, eng0, Nk = sp.nets[0], sp.halves[0].engine, N // 2           # per-kernel timing: one half
    else:
        net = DeviceNet(S, N, weights, mode=mode)
        sp = SelfPlay(None, n_games=N, net=net, training=True, seed=0, game_id_base=rank * N, use_graph=not a.no_graph,
                      board_size=S, simulation_per_step=sims, upper_simulation_per_step=upper)
        eng0, Nk = sp.engine, N
    stride = eng0.record_stride
    rec_cap = N * 4                                          # records exchanged per step (plies finishing per step ~ N)
    gather_out = torch.empty((world, rec_cap, stride), dtype=torch.uint8, device="cuda") if world > 1 else None
    gather_cnt = torch.zeros(world, dtype=torch.int64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev_nn = []

    def one_step(timed):
        """`sims` passes; NN forward time sampled with CUDA events every 25th pass (eager)."""
        sp.run_passes(sims)
        buf, games = sp.harvest()
        if world > 1:                                        # fill every rank's replay shard (NVLink allgather)
            cnt = torch.tensor([min(buf.shape[0], rec_cap)], dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(gather_cnt, cnt)
            dist.all_gather_into_tensor(gather_out.view(-1), sp.record_buf[:rec_cap].reshape(-1))
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
    for _ in range(a.warmup):
        one_step(False)
    c0 = sp.counters()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    recs = 0
    for _ in range(a.steps):
        r, _ = one_step(True)
        recs += r
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    c1 = sp.counters()
    clocks = sampler.stop() if rank == 0 else None

    # per-kernel-class timing inside the same workload, CUDA events on the launch stream(s).  Pipelined: one
    # half batch, block convs on the big partition, everything else on the small one -- as in the timed run.
    reps = 20
    if pipelined:
        sp.pipe.drain()
    torch.cuda.synchronize()
    if pipelined:
        prob, val = sp.pipe.prob[0], sp.pipe.value[0]
        fwd = lambda parts: check(lib.a5_net_forward_parts(net.handle, ct.c_void_p(eng0.planes_ptr), Nk, ptr(prob), ptr(val),
                                                           parts, stream_ptr()))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        body_ms = small_ms = tree_ms = 0.0
        for _ in range(reps):
            with torch.cuda.stream(part.small):
                ev[0].record(); fwd(1); ev[1].record()
            part.big.wait_stream(part.small)
            with torch.cuda.stream(part.big):
                ev[2].record(); fwd(2); ev[3].record()
            part.small.wait_stream(part.big)
            with torch.cuda.stream(part.small):
                fwd(4); ev[4].record(); eng0.step(prob, val); ev[5].record()
            torch.cuda.synchronize()
            body_ms += ev[2].elapsed_time(ev[3])
            small_ms += ev[0].elapsed_time(ev[1]) + ev[3].elapsed_time(ev[4])
            tree_ms += ev[4].elapsed_time(ev[5])
        body_ms /= reps; small_ms /= reps; tree_ms /= reps
        nn_ms = body_ms + small_ms
        c2 = sp.counters()
        with torch.cuda.stream(part.big):
            lt = layer_times(net, eng0.planes_ptr, Nk, prob, val)
        torch.cuda.synchronize()
    else:
        prob, val = sp.prob, sp.value
        tn0, tn1, tt0, tt1 = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        nn_ms = tree_ms = 0.0
        for _ in range(reps):
            tn0.record(); net.forward_raw(eng0.planes_ptr, N, prob, val); tn1.record()
            tt0.record(); eng0.step(prob, val); tt1.record()
            torch.cuda.synchronize()
            nn_ms += tn0.elapsed_time(tn1); tree_ms += tt0.elapsed_time(tt1)
        nn_ms /= reps; tree_ms /= reps
        c2 = sp.counters()
        lt = layer_times(net, eng0.planes_ptr, N, prob, val) if mode == _lib.NET_TC else None

    t = torch.tensor([ms, float(c1["moves"] - c0["moves"]), float(c1["leaf_evals"] - c0["leaf_evals"]),
                      float(c1["sims"] - c0["sims"]), float(c1["games"] - c0["games"])], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, moves, evals, nsims, games = tmax[0].item(), tsum[1].item(), tsum[2].item(), tsum[3].item(), tsum[4].item()
    else:
        ms, moves, evals, nsims, games = [x.item() for x in t]
    value = moves / (ms / 1000)

    # ---- end to end through the host-buffer API (BatchedPlayer.get_actions) ----------------
    import numpy as np
    del sp
    torch.cuda.empty_cache()
    if pipelined:
        bp = PipelinedBatchedPlayer(None, n_players=N, weights=weights, training=True, seed=1, game_id_base=rank * N,
                                    small_sms=a.small_sms, board_size=S, simulation_per_step=sims,
                                    upper_simulation_per_step=upper)
    else:
        bp = BatchedPlayer(None, n_players=N, net=net, training=True, seed=1, game_id_base=rank * N,
                           board_size=S, simulation_per_step=sims, upper_simulation_per_step=upper)
    boards = np.zeros((N, S, S), np.int8); last = np.full(N, -1, np.int32)
    clear = np.ones(N, np.uint8)
    e2e_moves, e2e_t = 0, 0.0
    for k in range(1 + a.e2e_steps):                          # first call is warm-up
        barrier()
        t0 = time.perf_counter()
        pol, act, nxt, codes = bp.get_actions(boards, last, None, clear, advance=True)
        over = codes != 0
        boards = np.where(over[:, None, None], 0, nxt).astype(np.int8)
        last = np.where(over, -1, act).astype(np.int32)
        clear = over.astype(np.uint8)
        barrier()
        if k > 0:
            e2e_t += time.perf_counter() - t0
            e2e_moves += N
    e2e = torch.tensor([e2e_t, float(e2e_moves)], dtype=torch.float64, device="cuda")
    if world > 1:
        em = e2e.clone(); dist.all_reduce(em, op=dist.ReduceOp.MAX)
        es = e2e.clone(); dist.all_reduce(es, op=dist.ReduceOp.SUM)
        e2e_val = es[1].item() / em[0].item() if em[0].item() > 0 else None
    else:
        e2e_val = e2e_moves / e2e_t if e2e_t > 0 else None

    if rank == 0:
        pk = peaks()
        flop = FLOP_PER_LEAF.get(S, FLOP_PER_LEAF[11] * S * S / 121)
        nn_tflops = flop * Nk / (nn_ms / 1000) / 1e12
        dsel = max(1, c2["sims"] - c1["sims"])
        # SURVEY 8(d) algorithmic bytes per simulation with the measured d, A, f_leaf of this run
        C = S * S
        dbar = (c1["selects"] - c0["selects"]) / max(1.0, c1["sims"] - c0["sims"])
        abar = (c1["legal_sum"] - c0["legal_sum"]) / max(1.0, c1["leaf_evals"] - c0["leaf_evals"])
        fleaf = (c1["leaf_evals"] - c0["leaf_evals"]) / max(1.0, c1["sims"] - c0["sims"])
        bytes_per_sim = dbar * (12 * abar + 16 + C) + dbar * 16 + dbar * 20 + fleaf * (3 * C + 4 * C + 4 + 12 * abar + C + 16)
        tree_gbs = bytes_per_sim * dsel / reps / (tree_ms / 1000) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": "moves/s", "leaf_evals_per_s": evals / (ms / 1000),
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if mode == _lib.NET_FP32 else "f16x2-split (fp32 accumulate)", "data": "synthetic",
            "config": {"workload": workload_name(a, world), "board": S, "sims": sims, "upper_sims": upper,
                       "games_per_gpu": N, "net_mode": "fp32" if mode == _lib.NET_FP32 else "tc",
                       "l2": "per-pass working set (activations + node arenas) exceeds the 126 MB L2",
                       "cuda_graph": (not a.no_graph) and not pipelined, "step": f"{sims} lock-step passes",
                       "schedule": (f"two half batches of {Nk} games pipelined on SM-partitioned streams (CUDA green contexts: "
                                    f"{part.n_big} SMs block convs, {part.n_small} SMs heads / tree pass / conv1)") if pipelined
                                   else "single stream, CUDA-graph replay of one pass",
                       "preroll": f"{a.preroll_moves} untimed moves at {a.preroll_sims} sims to spread games over all plies"},
            "e2e": {"value": e2e_val, "unit": "moves/s", "h2d_bytes_per_step": bp.h2d_bytes,
                    "d2h_bytes_per_step": bp.d2h_bytes, "api": ("PipelinedBatchedPlayer" if pipelined else "BatchedPlayer") + ".get_actions(host boards) + device step/terminal"},
            "gpu_launches": int((c1["passes"] - c0["passes"]) * (launches_per_pass(mode) + 1) * (2 if pipelined else 1)),
            "roofline": roofline_entry(mode, lt, nn_ms, nn_tflops, flop, Nk, C, pk,
                                       part=(part.n_big, part.n_small) if pipelined else None),
            "roofline_tree": {"bound": "hbm", "kernel": "k_step (tree pass)", "achieved": tree_gbs, "peak": pk["hbm"],
                              "unit": "GB/s", "frac": tree_gbs / pk["hbm"], "ms_per_launch": tree_ms,
                              "bytes_per_sim": bytes_per_sim, "d_bar": dbar, "a_bar": abar, "f_leaf": fleaf},
            "moves": moves, "sims_run": nsims, "games_finished": games, "records_per_step": recs / max(1, a.steps),
            "clocks": clocks,
        }
        if not a.no_cpu_baseline and world == 1:      # the CPU arm is reported at N = 1 only
            cores = max(1, min((os.cpu_count() or 1) - 1, 64))
            v, ev, wall = cpu_moves_per_sec(S, sims, upper, cores, 2)
            v5, ev5, _ = cpu_moves_per_sec(S, sims, upper, min(5, cores), 2)     # config.py:20 max_processes = 5
            out["cpu_baseline"] = {"value": v, "unit": "moves/s", "cores": cores, "kind": "port",
                                   "leaf_evals_per_s": ev, "five_workers": {"value": v5, "leaf_evals_per_s": ev5},
                                   "sample": f"{cores} processes x 2 moves of {sims} sims from the empty board, oracle port "
                                             f"(numpy MCTS + torch-CPU fp32 net, 1 thread each), {wall:.1f}s wall"}
        emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def roofline_entry(mode, lt, nn_ms, nn_tflops, flop, N, C, pk, part=None):
    """Dominant kernel = k_tc_conv2 (eight launches per pass -- block3/block4 conv1 run as one layer, and so
    do block3/block4 conv2 -- 98.9 % of the algorithmic FLOPs).  achieved = algorithmic conv FLOPs of one pass / summed
    CUDA-event time of those launches."""
    whole = {"kernel": "whole net forward (bitboards + conv1 + 8 block-conv launches + heads = 11 launches)", "achieved": nn_tflops, "frac": nn_tflops / pk["tensor"],
             "ms_per_pass": nn_ms, "flop_per_leaf": flop}
    if lt is None:
        return {"bound": "tensor", "kernel": "policy/value net forward, fp32 CUDA-core path", "achieved": nn_tflops,
                "peak": pk["tensor"], "unit": "TFLOP/s", "frac": nn_tflops / pk["tensor"],
                "peak_source": pk["src"] + " bf16 sustained", "traffic": None}
    conv_ms = sum(lt[1:11])
    conv_flop = 2.0 * CONV_MAC_PER_CELL * C * N
    ach = conv_flop / (conv_ms / 1000) / 1e12
    tr = measured_traffic()
    out_part = {}
    if part:
        out_part = {"boards_per_launch": N, "sms": part[0],
                    "note": f"launch group of one half batch ({N} boards) on the {part[0]}-SM partition while the other "
                            f"{part[1]} SMs run heads / tree pass / conv1 of the other half; peak is the whole chip's"}
    traffic = tr.get("conv_dram_bytes_per_pass") * N / 4096.0 if tr else None
    return {"bound": "tensor", "kernel": "k_tc_conv2 (tcgen05 cta_group::2 3x3 conv + residual, 8 launches per pass)",
            "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": ach / pk["tensor"],
            "peak_source": pk["src"] + " bf16 sustained (MEASURED_PEAKS.json)",
            "traffic": traffic, **out_part,
            "traffic_note": (tr.get("note") + f"; scaled to {N} boards") if tr else "no ncu capture committed for this build",
            "algorithmic_flop_per_pass": conv_flop, "issued_flop_factor": 3,
            "ms_per_pass": conv_ms, "ms_conv1": lt[0], "ms_heads": lt[11], "ms_layers": lt[1:11],
            "whole_net": whole}


def launches_per_pass(mode):
    # tensor-core path: bitboards + conv1 + 8 block-conv launches + fused dense heads; fp32 path: conv1 + 10 + 4 head kernels
    return 11 if mode == 1 else 15


CONV_MAC_PER_CELL = 485_376          # the ten 3x3(+1x1) block convs, MACs per board cell (58,730,496 @ 11x11)


def layer_times(net, planes_ptr, N, prob, val, reps=10):
    """CUDA-event time of every launch group of one net forward (ms): conv1, 10 block convs, heads."""
    import ctypes as C
    from alphafive_b200 import _lib
    from alphafive_b200._lib import check, ptr, stream_ptr
    fn = _lib.load().a5__debug_layer_times
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_void_p]
    ms = (C.c_float * 12)()
    check(fn(net.handle, C.c_void_p(planes_ptr), N, reps, ptr(prob), ptr(val), ms, stream_ptr()))
    return [float(x) for x in ms]


def measured_traffic():
    """DRAM bytes per pass of the block convs from the committed ncu capture (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


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
