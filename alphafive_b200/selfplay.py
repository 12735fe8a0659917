"""Lock-step self-play drivers.

``SelfPlay``       -- the data-generating loop of the reference (main.py:82-94 ->
                      Player.run, player.py:53-82) for N games at once, entirely on the
                      device: games are played, recorded and restarted by the tree kernel;
                      the host only launches passes and harvests finished-game records.
``BatchedPlayer``  -- N ``Player`` objects behind one call with HOST buffers:
                      ``get_actions(boards, last_actions) -> (policies, actions)`` is
                      ``Player.get_action`` (player.py:128-147) batched; host<->device
                      copies go through pinned memory.  This is the reference-facing API
                      the end-to-end number of bench.py is measured through.
"""
from __future__ import annotations

import numpy as np
import torch

from . import rules
from .engine import SearchEngine, make_config
from .replay import gather_records, parse_records, records_to_games
from .net import DeviceNet


class EvalCache:
    """Cross-game evaluation cache (a5_evalcache_*): leaves whose position some game of the batch has had evaluated
    before are served from a device table; the others go through the network as a compact batch of ``cap`` boards.
    ``cap`` defaults to the batch that fills whole waves of the persistent conv kernels at the expected demand."""

    def __init__(self, S, n_games, log2_slots=21, cap=None, num_sms=None):
        import ctypes as C
        from . import _lib
        from ._lib import check
        self.lib = _lib.load()
        if cap is None:
            # Whole waves: the persistent conv kernels run ceil(groups / CTAs) waves of 256-position groups.  About 93 %
            # of the games ask for a network evaluation in a pass (2 % of the simulations end in terminal positions,
            # ~5 % of the leaves are table hits); the compact batch is the largest one that fills its waves at that
            # demand -- 14 waves = 3683 boards for 4096 games of 11x11 -- and the few leaves beyond it wait a pass.
            # (Measured at 4096 x 11x11: 3683 -> 5,470 moves/s, 3946 (15 waves) -> 5,406, 3420 (13) -> 5,378.)
            sms = num_sms or torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
            per_board, ctas = (S + 1) * (S + 1), sms & ~1
            waves = int(0.93 * n_games * per_board / 256 / ctas)
            cap = min(n_games, (waves * ctas * 256) // per_board) if waves >= 4 else n_games
        self.S, self.N, self.cap = S, n_games, int(cap)
        h = C.c_void_p()
        check(self.lib.a5_evalcache_create(S, n_games, log2_slots, self.cap, C.byref(h)))
        self.handle = h
        self.planes_ptr = self.lib.a5_evalcache_planes(h)
        dev = torch.device("cuda", torch.cuda.current_device())
        from .engine import _view
        self.cprob = _view(self.lib.a5_evalcache_prob(h), (self.cap, S * S), torch.float32, dev)
        self.cvalue = _view(self.lib.a5_evalcache_value(h), (self.cap,), torch.float32, dev)
        self.served = torch.ones((n_games,), dtype=torch.uint8, device=dev)
        self.net_version = None

    def clear(self):
        from ._lib import check, stream_ptr
        check(self.lib.a5_evalcache_clear(self.handle, stream_ptr()))

    def stats(self) -> dict:
        import ctypes as C
        from ._lib import check, stream_ptr
        arr = (C.c_int64 * 4)()
        check(self.lib.a5_evalcache_stats(self.handle, arr, stream_ptr()))
        return dict(lookups=arr[0], hits=arr[1], deferred=arr[2], stored=arr[3])

    def close(self):
        if self.handle:
            self.lib.a5_evalcache_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SelfPlay:
    def __init__(self, cfg=None, n_games=4096, net: DeviceNet | None = None, training=True, seed=0,
                 game_id_base=0, use_graph=True, _defer_net=False, eval_cache=False, **cfg_kw):
        self.config = make_config(cfg, n_games=n_games, training=training, auto_play=True, seed=seed,
                                  game_id_base=game_id_base, **cfg_kw)
        self.engine = SearchEngine(self.config)
        self.N, self.S = n_games, self.config.board_size
        self.net = net if (net is not None or _defer_net) else DeviceNet(self.S, n_games)
        dev = self.engine.device
        self.prob = torch.zeros((n_games, self.S * self.S), dtype=torch.float32, device=dev)
        self.value = torch.zeros((n_games,), dtype=torch.float32, device=dev)
        self.record_buf = torch.empty((self.config.record_capacity or n_games * self.S * self.S,
                                       self.engine.record_stride), dtype=torch.uint8, device=dev)
        self.use_graph = use_graph
        self._graph = None
        self._started = False
        self.passes = 0
        # eval_cache: True, or dict(log2_slots=..., cap=...) -- see EvalCache
        self.cache = None
        if eval_cache:
            self.cache = EvalCache(self.S, n_games, **(eval_cache if isinstance(eval_cache, dict) else {}))

    # one pass = network forward over all N pending leaves + one tree-kernel pass
    def _pass(self):
        if self.cache is None:
            self.net.forward_raw(self.engine.planes_ptr, self.N, self.prob, self.value)
            self.engine.step(self.prob, self.value)
            return
        self.engine.cached_pass(self.net, self.cache, self.prob, self.value)

    def _check_cache(self):
        """Cached results belong to one set of weights."""
        if self.cache is not None and self.cache.net_version != self.net.version:
            self.cache.clear()
            self.cache.net_version = self.net.version

    def start(self):
        self._check_cache()
        if not self._started:
            self.engine.step()              # first descent: every game reaches its (unseen) root
            self._started = True
            if self.use_graph:
                # warm up on a side stream, then capture one pass
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._pass()
                torch.cuda.current_stream().wait_stream(s)
                self.passes += 1
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):     # capture only: nothing executes here
                    self._pass()

    def set_budget(self, sims: int, upper: int):
        """config.simulation_per_step / upper_simulation_per_step for the moves that start from
        now on (the reference reads them lazily per get_action).  Kernel parameters are baked
        into the captured graph, so the pass is re-captured."""
        self.engine.set_budget(sims, upper)
        if self._graph is not None:
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._pass()

    def run_passes(self, k: int):
        self._check_cache()
        self.start()
        if self._graph is not None:
            for _ in range(k):
                self._graph.replay()
        else:
            for _ in range(k):
                self._pass()
        self.passes += k

    def harvest(self, buf=None):
        """(records uint8 [count, stride] on the device, games finished since last call)."""
        return self.engine.harvest(self.record_buf if buf is None else buf)

    def kernel_accounting(self, passes=200):
        """In-situ time of every kernel of a pass, measured inside CUDA-graph replays of the running workload
        (%globaltimer stamps, a5__debug_ktime_*): dict name -> (us from the predecessor's end to this
        kernel's end, us from its first CTA's start to its end); the first values partition the pass exactly.
        The games advance by ``passes + 1`` passes."""
        import ctypes as C
        from . import _lib
        from ._lib import check, stream_ptr
        lib = _lib.load()
        for name in ("a5__debug_ktime_enable", "a5__debug_ktime_fold", "a5__debug_ktime_read"):
            getattr(lib, name).restype = C.c_int
        lib.a5__debug_ktime_fold.argtypes = [C.c_void_p]
        lib.a5__debug_ktime_read.argtypes = [C.POINTER(C.c_double)]
        self.start()
        torch.cuda.synchronize()
        check(lib.a5__debug_ktime_enable(1))
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._pass()
                check(lib.a5__debug_ktime_fold(stream_ptr()))
            for _ in range(passes + 1):            # the first replay only sets the reference point
                g.replay()
            out = (C.c_double * 48)()
            check(lib.a5__debug_ktime_read(out))
        finally:
            check(lib.a5__debug_ktime_enable(0))
        self.passes += passes + 1
        names = ["k_c1_bits", "k_tc_conv1m"] + [f"k_tc_conv2[{i}]" for i in range(8)] + ["k_tc_fc", "k_step", "(fold)",
                                                                                          "k_ec_lookup", "k_ec_commit"]
        table = {}
        for i in [13] + list(range(11)) + [14, 11, 12]:          # slots in the order their kernels run
            cnt = out[3 * i + 2]
            if cnt > 0:
                table[names[i]] = (out[3 * i] / cnt / 1000.0, out[3 * i + 1] / cnt / 1000.0)
        if "k_tc_conv2[0]" in table and "k_tc_conv2[1]" not in table:      # one launch for all block convs
            table = {("k_tc_mega" if k == "k_tc_conv2[0]" else k): v for k, v in table.items()}
        return table

    def harvest_games(self):
        """Finished games as the reference's replay tuples (player.py:77-82):
        list of (record list, result) like main.py:94's q.put((game_record, result))."""
        buf, _ = self.harvest()
        return records_to_games(parse_records(buf, self.S), self.S)

    def harvest_all_ranks(self, group=None):
        """Harvest this rank's finished-ply records and gather every rank's (NCCL allgather over
        NVLink; the one collective of the path).  Returns (records uint8 [total, stride], counts)."""
        buf, _ = self.harvest()
        return gather_records(buf, group=group)

    def counters(self):
        return self.engine.counters()


class PipelinedSelfPlay:
    """``SelfPlay`` as two half batches on SM-partitioned streams (alphafive_b200.pipeline): the same games,
    records and counters -- game g of the first half is global game ``game_id_base + g``, of the second
    ``game_id_base + n_games / 2 + g`` -- with heads / tree pass / conv1 of one half hidden behind the block
    convs of the other."""

    def __init__(self, cfg=None, n_games=4096, weights=None, training=True, seed=0, game_id_base=0, small_sms=16,
                 **cfg_kw):
        from .net import glorot_init
        from .pipeline import SmPartition, TwoHalfPipeline
        assert n_games % 2 == 0
        self.N, h = n_games, n_games // 2
        cfg_kw.setdefault("max_inner", 2 if n_games >= 256 else 16)      # the engine's default for the whole batch
        self.halves = [SelfPlay(cfg, n_games=h, net=None, training=training, seed=seed, game_id_base=game_id_base + i * h,
                                use_graph=False, _defer_net=True, **cfg_kw) for i in (0, 1)]
        self.config, self.S = self.halves[0].config, self.halves[0].S
        w = weights if weights is not None else glorot_init(self.S, 0)
        self.nets = [DeviceNet(self.S, h, w) for _ in (0, 1)]
        for sp, net in zip(self.halves, self.nets):
            sp.net = net
        self.pipe = TwoHalfPipeline([sp.engine for sp in self.halves], self.nets, SmPartition.get(small_sms))
        self.record_buf = torch.empty((sum(sp.record_buf.shape[0] for sp in self.halves), self.halves[0].engine.record_stride),
                                      dtype=torch.uint8, device=self.halves[0].engine.device)
        self.passes = 0

    def set_weights(self, weights):
        self.pipe.drain()
        torch.cuda.current_stream().synchronize()
        for n in self.nets:
            n.set_weights(weights)

    def start(self):
        if not self.pipe.primed:
            self.pipe.prime()

    def set_budget(self, sims: int, upper: int):
        self.pipe.drain()
        for sp in self.halves:
            sp.engine.set_budget(sims, upper)
        for st in (self.pipe.part.small, self.pipe.part.big):
            st.wait_stream(torch.cuda.current_stream())

    def run_passes(self, k: int):
        self.start()
        self.pipe.run(k)
        self.passes += k

    def harvest(self):
        self.pipe.drain()
        n = g = 0
        for sp in self.halves:
            buf, games = sp.engine.harvest(sp.record_buf)
            self.record_buf[n:n + buf.shape[0]] = buf
            n += buf.shape[0]
            g += games
        for st in (self.pipe.part.small, self.pipe.part.big):       # the next passes must see the reset arenas
            st.wait_stream(torch.cuda.current_stream())
        return self.record_buf[:n], g

    def harvest_games(self):
        buf, _ = self.harvest()
        return records_to_games(parse_records(buf, self.S), self.S)

    def harvest_all_ranks(self, group=None):
        buf, _ = self.harvest()
        return gather_records(buf, group=group)

    def counters(self):
        self.pipe.drain()
        torch.cuda.current_stream().synchronize()
        a, b = (sp.engine.counters() for sp in self.halves)
        return {k: (max(a[k], b[k]) if k in ("max_nodes", "passes") else a[k] + b[k]) for k in a}


class BatchedPlayer:
    """N reference ``Player`` objects in lock-step, host buffers in and out."""

    def __init__(self, cfg=None, n_players=1, net: DeviceNet | None = None, training=True, random_a=False,
                 seed=0, game_id_base=0, check_every=16, eval_cache=False, **cfg_kw):
        self.config = make_config(cfg, n_games=n_players, training=training, random_a=random_a,
                                  auto_play=False, seed=seed, game_id_base=game_id_base, **cfg_kw)
        self.engine = SearchEngine(self.config)
        self.N, self.S = n_players, self.config.board_size
        self.C = self.S * self.S
        self.net = net if net is not None else DeviceNet(self.S, n_players)
        self.check_every = check_every
        # eval_cache: True, dict(log2_slots=..., cap=...) or an EvalCache to share (see EvalCache)
        self.cache = eval_cache if isinstance(eval_cache, EvalCache) else (
            EvalCache(self.S, n_players, **(eval_cache if isinstance(eval_cache, dict) else {})) if eval_cache else None)
        dev = self.engine.device
        N, S, C = self.N, self.S, self.C
        pin = lambda *shape, dtype: torch.empty(shape, dtype=dtype).pin_memory()
        self.h_boards, self.h_last = pin(N, S, S, dtype=torch.int8), pin(N, dtype=torch.int32)
        self.h_policy, self.h_action = pin(N, C, dtype=torch.float32), pin(N, dtype=torch.int32)
        self.h_next, self.h_codes = pin(N, S, S, dtype=torch.int8), pin(N, dtype=torch.int8)
        self.d_boards = torch.empty((N, S, S), dtype=torch.int8, device=dev)
        self.d_last = torch.empty((N,), dtype=torch.int32, device=dev)
        self.h2d_bytes = N * C + 4 * N
        self.d2h_bytes = N * C * 4 + 4 * N + N * C + N

    def get_actions(self, boards, last, active=None, clear=None, advance=False):
        """boards int8[N,S,S] / last int32[N] host arrays -> (policy f32[N,S,S], action int32[N])
        host arrays.  With ``advance`` the played positions and their terminal codes are
        returned as well (utils.step + utils.is_game_over on the device)."""
        self.h_boards.copy_(torch.as_tensor(boards, dtype=torch.int8).reshape(self.N, self.S, self.S))
        self.h_last.copy_(torch.as_tensor(last, dtype=torch.int32))
        self.d_boards.copy_(self.h_boards, non_blocking=True)
        self.d_last.copy_(self.h_last, non_blocking=True)
        self.engine.set_roots(self.d_boards, self.d_last, active, clear)
        self.engine.run_search(net=self.net, check_every=self.check_every, cache=self.cache)
        policy, action = self.engine.finish_move()
        self.h_policy.copy_(policy, non_blocking=True)
        self.h_action.copy_(action, non_blocking=True)
        if advance:
            nxt = rules.step(self.d_boards, action.clamp(min=0))
            codes = rules.terminal(nxt, self.config.goal)
            self.h_next.copy_(nxt, non_blocking=True)
            self.h_codes.copy_(codes, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        pol = self.h_policy.numpy().reshape(self.N, self.S, self.S)
        if advance:
            return pol, self.h_action.numpy(), self.h_next.numpy(), self.h_codes.numpy()
        return pol, self.h_action.numpy()


    # ---- continuous batching: searches are collected as they end, the other players keep searching -----------
    def start_stream(self, boards, last, clear=None, passes=4, cap=None):
        """Continuous form of ``get_actions``.  The searches of a batch end in different passes (the budget rule of
        player.py:140-143 cuts re-used trees short, simulations that end in terminal positions run two to a pass, leaves
        deferred by the evaluation cache wait a pass); ``get_actions`` returns when the slowest is done, the
        reference's players are independent objects and never wait for each other.  Here every player gets
        its root (host arrays as in ``get_actions``), then the caller alternates

            games, policy, action, nxt, codes = bp.poll()      # searches that have ended (host arrays)
            bp.submit(games, next_boards, last, clear)         # their next roots (host arrays)

        ``poll`` queues ``passes`` search passes before it waits for its results, so the device never idles while
        the host decides; a collected player is parked until its next root arrives.  Per-player results are those
        of ``get_actions`` (every game owns its table and its counter-based RNG stream)."""
        N, S, C = self.N, self.S, self.C
        dev = self.engine.device
        cap = int(cap or max(64, N // 32))       # rows per poll: ~4 x the mean arrivals at 4 passes per poll and ~500 sims
        pin = lambda *shape, dtype: torch.empty(shape, dtype=dtype).pin_memory()
        dv = lambda *shape, dtype: torch.empty(shape, dtype=dtype, device=dev)
        self._s = dict(
            cap=cap, passes=int(passes), ev=torch.cuda.Event(), sub_ev=None,
            d_count=dv(1, dtype=torch.int32), d_game=dv(cap, dtype=torch.int32), d_policy=dv(cap, C, dtype=torch.float32),
            d_action=dv(cap, dtype=torch.int32), d_next=dv(cap, C, dtype=torch.int8), d_code=dv(cap, dtype=torch.int8),
            h_count=pin(1, dtype=torch.int32), h_game=pin(cap, dtype=torch.int32), h_policy=pin(cap, C, dtype=torch.float32),
            h_action=pin(cap, dtype=torch.int32), h_next=pin(cap, C, dtype=torch.int8), h_code=pin(cap, dtype=torch.int8),
            s_game=pin(N, dtype=torch.int32), s_boards=pin(N, C, dtype=torch.int8), s_last=pin(N, dtype=torch.int32),
            s_clear=pin(N, dtype=torch.uint8),
            sd_game=dv(N, dtype=torch.int32), sd_boards=dv(N, C, dtype=torch.int8), sd_last=dv(N, dtype=torch.int32),
            sd_clear=dv(N, dtype=torch.uint8), h2d=0, d2h=0, polls=0)
        self.h_boards.copy_(torch.as_tensor(boards, dtype=torch.int8).reshape(N, S, S))
        self.h_last.copy_(torch.as_tensor(last, dtype=torch.int32))
        self.d_boards.copy_(self.h_boards, non_blocking=True)
        self.d_last.copy_(self.h_last, non_blocking=True)
        self._s["h2d"] += N * C + 4 * N
        self.engine.set_roots(self.d_boards, self.d_last, None, clear)
        self.engine.step()                                   # first descent: every game reaches its root

    def poll(self):
        """-> (games int32[n], policy f32[n,S,S], action int32[n], next int8[n,S,S], codes int8[n]): the searches
        that have ended since the last poll (at most ``cap`` of them; the rest comes with the next one)."""
        s = self._s
        self.engine.collect_moves(s["cap"], s["d_count"], s["d_game"], s["d_policy"], s["d_action"], s["d_next"], s["d_code"])
        for k in ("count", "game", "policy", "action", "next", "code"):
            s["h_" + k].copy_(s["d_" + k], non_blocking=True)
        s["ev"].record()
        self.engine.replay_passes(s["passes"], self.net, self.cache)     # the device works on while the host waits / decides
        s["ev"].synchronize()
        n = min(int(s["h_count"][0]), s["cap"])
        s["d2h"] += sum(s["h_" + k].numel() * s["h_" + k].element_size() for k in ("count", "game", "policy", "action", "next", "code"))
        s["polls"] += 1
        return (s["h_game"][:n].numpy().copy(), s["h_policy"][:n].numpy().reshape(n, self.S, self.S).copy(),
                s["h_action"][:n].numpy().copy(), s["h_next"][:n].numpy().reshape(n, self.S, self.S).copy(),
                s["h_code"][:n].numpy().copy())

    def stream_stats(self) -> dict:
        """Bytes copied host->device / device->host and polls since ``start_stream`` (what bench.py reports)."""
        s = self._s
        return dict(h2d_bytes=s["h2d"], d2h_bytes=s["d2h"], polls=s["polls"], collect_cap=s["cap"], passes_per_poll=s["passes"])

    def submit(self, games, boards, last, clear=None):
        """Next roots of the players ``games`` (host arrays: boards int8[n,S,S], last int32[n], clear uint8[n])."""
        s = self._s
        n = len(games)
        if n == 0:
            return
        if s["sub_ev"] is not None:
            s["sub_ev"].synchronize()                        # the staging buffers of the previous submit are free again
        s["s_game"][:n].copy_(torch.as_tensor(games, dtype=torch.int32))
        s["s_boards"][:n].copy_(torch.as_tensor(boards, dtype=torch.int8).reshape(n, self.C))
        s["s_last"][:n].copy_(torch.as_tensor(last, dtype=torch.int32))
        if clear is None:
            s["s_clear"][:n].zero_()
        else:
            s["s_clear"][:n].copy_(torch.as_tensor(clear, dtype=torch.uint8))
        for k in ("game", "boards", "last", "clear"):
            s["sd_" + k][:n].copy_(s["s_" + k][:n], non_blocking=True)
        s["h2d"] += n * (4 + self.C + 4 + 1)
        self.engine.submit_roots(n, s["sd_game"], s["sd_boards"], s["sd_last"], s["sd_clear"])
        s["sub_ev"] = torch.cuda.Event()
        s["sub_ev"].record()


class PipelinedBatchedPlayer(BatchedPlayer):
    """``BatchedPlayer`` with the search of the two halves of the batch pipelined on SM-partitioned streams
    (alphafive_b200.pipeline); same call, same host buffers, same results."""

    def __init__(self, cfg=None, n_players=2, weights=None, training=True, random_a=False, seed=0, game_id_base=0,
                 small_sms=16, **cfg_kw):
        from .net import glorot_init
        from .pipeline import SmPartition, TwoHalfPipeline
        assert n_players % 2 == 0
        cfg_kw.setdefault("max_inner", 2 if n_players >= 256 else 16)
        h = n_players // 2
        self.config = make_config(cfg, n_games=h, training=training, random_a=random_a, auto_play=False, seed=seed,
                                  game_id_base=game_id_base, **cfg_kw)
        self.N, self.S = n_players, self.config.board_size
        self.C = self.S * self.S
        self.engines = [SearchEngine(make_config(cfg, n_games=h, training=training, random_a=random_a, auto_play=False,
                                                 seed=seed, game_id_base=game_id_base + i * h, **cfg_kw)) for i in (0, 1)]
        w = weights if weights is not None else glorot_init(self.S, 0)
        self.nets = [DeviceNet(self.S, h, w) for _ in (0, 1)]
        self.pipe = TwoHalfPipeline(self.engines, self.nets, SmPartition.get(small_sms))
        dev = self.engines[0].device
        N, S, C = self.N, self.S, self.C
        pin = lambda *shape, dtype: torch.empty(shape, dtype=dtype).pin_memory()
        self.h_boards, self.h_last = pin(N, S, S, dtype=torch.int8), pin(N, dtype=torch.int32)
        self.h_policy, self.h_action = pin(N, C, dtype=torch.float32), pin(N, dtype=torch.int32)
        self.h_next, self.h_codes = pin(N, S, S, dtype=torch.int8), pin(N, dtype=torch.int8)
        self.d_boards = torch.empty((N, S, S), dtype=torch.int8, device=dev)
        self.d_last = torch.empty((N,), dtype=torch.int32, device=dev)
        self.h2d_bytes = N * C + 4 * N
        self.d2h_bytes = N * C * 4 + 4 * N + N * C + N

    def get_actions(self, boards, last, active=None, clear=None, advance=False):
        N, h = self.N, self.N // 2
        self.h_boards.copy_(torch.as_tensor(boards, dtype=torch.int8).reshape(N, self.S, self.S))
        self.h_last.copy_(torch.as_tensor(last, dtype=torch.int32))
        self.d_boards.copy_(self.h_boards, non_blocking=True)
        self.d_last.copy_(self.h_last, non_blocking=True)
        half = lambda a, i: None if a is None else np.asarray(a)[i * h:(i + 1) * h]
        for i, e in enumerate(self.engines):
            e.set_roots(self.d_boards[i * h:(i + 1) * h], self.d_last[i * h:(i + 1) * h], half(active, i), half(clear, i))
        left = max(1, max(int(e.sims_left().max().item()) for e in self.engines))
        self.pipe.prime()
        while True:
            self.pipe.run(left)
            self.pipe.drain()
            if sum(e.busy() for e in self.engines) == 0:
                break
            left = max(1, max(int(e.sims_left().max().item()) for e in self.engines))
        outs = [e.finish_move() for e in self.engines]
        policy, action = torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])
        self.h_policy.copy_(policy, non_blocking=True)
        self.h_action.copy_(action, non_blocking=True)
        if advance:
            nxt = rules.step(self.d_boards, action.clamp(min=0))
            codes = rules.terminal(nxt, self.config.goal)
            self.h_next.copy_(nxt, non_blocking=True)
            self.h_codes.copy_(codes, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        pol = self.h_policy.numpy().reshape(N, self.S, self.S)
        if advance:
            return pol, self.h_action.numpy(), self.h_next.numpy(), self.h_codes.numpy()
        return pol, self.h_action.numpy()
