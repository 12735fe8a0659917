"""SearchEngine: N concurrent MCTS games on the device (a5_engine_*).

Python mirror of the reference's search loop for many games at once: the object plays
the role of N ``genData.player.Player`` instances stepping in lock-step."""
from __future__ import annotations

import os
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import Config, RecordHeader, check, ptr, stream_ptr

COUNTER_NAMES = ("moves", "sims", "leaf_evals", "terminal_leaves", "selects", "legal_sum", "games",
                 "max_nodes", "overflows", "passes", "records_dropped")


def make_config(cfg=None, **kw) -> Config:
    """a5_config from a reference-style config module/object (config.py attribute names)."""
    g = lambda name, default: kw.pop(name, getattr(cfg, name, default) if cfg is not None else default)
    c = Config()
    c.board_size = g("board_size", 11)
    c.goal = g("goal", 5)
    c.sims = g("simulation_per_step", 542)
    c.upper_sims = g("upper_simulation_per_step", 642)
    c.c_puct = g("c_puct", 5.0)
    c.dirichlet_alpha = g("dirichlet_alpha", 0.3)
    c.init_temp = g("init_temp", 1.2)
    c.tau_decay = g("tau_decay_rate", 0.94)
    c.tau_decay_r = g("tau_decay_rate_r", 0.9)
    c.gamma = g("gamma", 0.94)
    c.n_games = kw.pop("n_games", 1)
    c.training = int(kw.pop("training", True))
    c.random_a = int(kw.pop("random_a", False))
    c.auto_play = int(kw.pop("auto_play", False))
    c.node_capacity = kw.pop("node_capacity", 0)
    c.max_inner = kw.pop("max_inner", int(os.environ.get("A5_MAX_INNER", "0")))
    c.seed = kw.pop("seed", 0)
    c.game_id_base = kw.pop("game_id_base", 0)
    c.record_capacity = kw.pop("record_capacity", 0)
    assert not kw, f"unknown config keys {sorted(kw)}"
    return c


class SearchEngine:
    def __init__(self, config: Config, device=None):
        self.lib = _lib.load()
        self.cfg = config
        self.N, self.S = config.n_games, config.board_size
        self.C = self.S * self.S
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.a5_engine_create(C.byref(config), C.byref(h)))
        self.handle = h
        self.planes_ptr = self.lib.a5_engine_planes(h)
        self.record_stride = self.lib.a5_record_stride(self.S)
        self._first = True
        # CUDA graph of one search pass (leaf evaluation + tree pass) for run_search: kernel parameters are baked
        # into it, so it is keyed on the net, its compute path and a version that set_mode / set_budget bump
        self._version = 0
        self._mode = (bool(config.training), bool(config.random_a))
        self._graph = None
        self._graph_key = None
        self._prob = self._value = None
        self.use_graph = os.environ.get("A5_SEARCH_GRAPH", "1") != "0"

    # -- raw passes -----------------------------------------------------------
    def reset(self):
        check(self.lib.a5_engine_reset(self.handle, stream_ptr()))
        self._first = True

    def set_roots(self, boards, last=None, active=None, clear=None):
        """boards int8[N,S,S]; last int32[N] flat cell or -1; active/clear uint8[N]."""
        dev = self.device
        boards = torch.as_tensor(boards, dtype=torch.int8).to(dev).contiguous()
        last_t = None if last is None else torch.as_tensor(last, dtype=torch.int32).to(dev).contiguous()
        act_t = None if active is None else torch.as_tensor(active, dtype=torch.uint8).to(dev).contiguous()
        clr_t = None if clear is None else torch.as_tensor(clear, dtype=torch.uint8).to(dev).contiguous()
        check(self.lib.a5_engine_set_roots(self.handle, ptr(boards), ptr(last_t), ptr(act_t), ptr(clr_t), stream_ptr()))
        self._first = True

    def set_mode(self, training: bool, random_a: bool = False):
        check(self.lib.a5_engine_set_mode(self.handle, int(training), int(random_a)))
        if self._mode != (bool(training), bool(random_a)):       # Player.get_action sets it on every call
            self._mode = (bool(training), bool(random_a))
            self._version += 1

    def set_budget(self, sims: int, upper: int):
        check(self.lib.a5_engine_set_budget(self.handle, int(sims), int(upper)))
        self._version += 1

    def step(self, prob=None, value=None):
        if self._first:
            prob = value = None
            self._first = False
        check(self.lib.a5_engine_step(self.handle, ptr(prob), ptr(value), stream_ptr()))

    def step_served(self, prob, value, served):
        """step() when an evaluation cache decided which pending leaves were evaluated this pass (served uint8 [N])."""
        check(self.lib.a5_engine_step_served(self.handle, ptr(prob), ptr(value), ptr(served), stream_ptr()))

    def cached_pass(self, net, cache, prob, value, net_mode=None):
        """One search pass through a cross-game evaluation cache (selfplay.EvalCache): table hits are copied to
        the games' rows, the misses go through the network as a compact batch, the rest waits a pass."""
        check(self.lib.a5_evalcache_lookup(cache.handle, self.planes_ptr, self.lib.a5_engine_need_eval(self.handle), ptr(prob),
                                           ptr(value), ptr(cache.served), stream_ptr()))
        net.forward_raw(cache.planes_ptr, cache.cap, cache.cprob, cache.cvalue, net_mode)
        check(self.lib.a5_evalcache_commit(cache.handle, ptr(prob), ptr(value), stream_ptr()))
        self.step_served(prob, value, cache.served)

    def planes(self) -> torch.Tensor:
        """Zero-copy int8 [N,3,S,S] view of the leaf planes written by the last pass."""
        return _view(self.planes_ptr, (self.N, 3, self.S, self.S), torch.int8, self.device)

    def need_eval(self) -> torch.Tensor:
        return _view(self.lib.a5_engine_need_eval(self.handle), (self.N,), torch.uint8, self.device)

    def sims_left(self) -> torch.Tensor:
        return _view(self.lib.a5_engine_sims_left(self.handle), (self.N,), torch.int32, self.device)

    def busy(self) -> int:
        b = C.c_int32()
        check(self.lib.a5_engine_busy(self.handle, C.byref(b), stream_ptr()))
        return b.value

    def finish_move(self):
        policy = torch.empty((self.N, self.C), dtype=torch.float32, device=self.device)
        action = torch.empty((self.N,), dtype=torch.int32, device=self.device)
        check(self.lib.a5_engine_finish_move(self.handle, ptr(policy), ptr(action), stream_ptr()))
        return policy, action

    def collect_moves(self, cap, count, game, policy, action, nxt, code):
        """a5_engine_collect_moves into caller-owned device buffers (see include/alphafive.h): every search that has
        ended since the last call -> compact rows (game, calc_policy row, action, next position, terminal code);
        those games are parked until submit_roots gives them their next root."""
        check(self.lib.a5_engine_collect_moves(self.handle, int(cap), ptr(count), ptr(game), ptr(policy), ptr(action),
                                               ptr(nxt), ptr(code), stream_ptr()))

    def submit_roots(self, n, game, boards, last=None, clear=None):
        """a5_engine_submit_roots: Player.get_action entry for the n games listed in ``game`` (device tensors);
        every other game is left as it is."""
        check(self.lib.a5_engine_submit_roots(self.handle, int(n), ptr(game), ptr(boards), ptr(last), ptr(clear), stream_ptr()))

    def root_stats(self):
        n = torch.empty((self.N, self.C), dtype=torch.int32, device=self.device)
        w = torch.empty((self.N, self.C), dtype=torch.float32, device=self.device)
        p = torch.empty((self.N, self.C), dtype=torch.float32, device=self.device)
        s = torch.empty((self.N,), dtype=torch.int32, device=self.device)
        check(self.lib.a5_engine_root_stats(self.handle, ptr(n), ptr(w), ptr(p), ptr(s), stream_ptr()))
        return n, w, p, s

    def node_stats(self, boards):
        """(n, w, p, sum_n) of the table node of ``boards[i]`` in game i (sum_n = -1 if absent)."""
        boards = torch.as_tensor(boards, dtype=torch.int8).to(self.device).contiguous()
        n = torch.empty((self.N, self.C), dtype=torch.int32, device=self.device)
        w = torch.empty((self.N, self.C), dtype=torch.float32, device=self.device)
        p = torch.empty((self.N, self.C), dtype=torch.float32, device=self.device)
        s = torch.empty((self.N,), dtype=torch.int32, device=self.device)
        check(self.lib.a5_engine_node_stats(self.handle, ptr(boards), ptr(n), ptr(w), ptr(p), ptr(s), stream_ptr()))
        return n, w, p, s

    def roots(self):
        """(boards int8[N,S,S], last int32[N]) -- where every game stands (copies)."""
        boards = torch.empty((self.N, self.S, self.S), dtype=torch.int8, device=self.device)
        last = torch.empty((self.N,), dtype=torch.int32, device=self.device)
        check(self.lib.a5_engine_get_roots(self.handle, ptr(boards), ptr(last), stream_ptr()))
        return boards, last

    def tau(self) -> torch.Tensor:
        """Zero-copy float64 [N] view of every game's temperature (Player.tau)."""
        return _view(self.lib.a5_engine_tau(self.handle), (self.N,), torch.float64, self.device)

    def table_dump(self, game: int, max_nodes: int = 65536):
        boards = torch.empty((max_nodes, self.S, self.S), dtype=torch.int8, device=self.device)
        sums = torch.empty((max_nodes,), dtype=torch.int32, device=self.device)
        cnt = C.c_int32()
        check(self.lib.a5_engine_table_dump(self.handle, game, ptr(boards), ptr(sums), max_nodes, C.byref(cnt), stream_ptr()))
        k = min(cnt.value, max_nodes)
        return boards[:k].cpu().numpy(), sums[:k].cpu().numpy()

    def counters(self) -> dict:
        arr = (C.c_int64 * _lib.NUM_COUNTERS)()
        check(self.lib.a5_engine_counters(self.handle, arr, stream_ptr()))
        return {k: int(arr[i]) for i, k in enumerate(COUNTER_NAMES)}

    def harvest(self, buf: torch.Tensor | None = None):
        """Finished-ply records as a uint8 [count, stride] CUDA tensor + games completed."""
        cap = self.cfg.record_capacity or self.N * self.C
        if buf is None:
            buf = torch.empty((cap, self.record_stride), dtype=torch.uint8, device=self.device)
        cnt, games = C.c_int32(), C.c_int32()
        check(self.lib.a5_engine_harvest(self.handle, ptr(buf), buf.shape[0], C.byref(cnt), C.byref(games), stream_ptr()))
        return buf[:cnt.value], games.value

    def replay_passes(self, k: int, net, cache=None, net_mode=None):
        """``k`` search passes back to back on the current stream (leaf evaluation by the on-device ``net`` +
        tree pass; through ``cache``, a selfplay.EvalCache, when given).  The pass is captured into a CUDA graph
        the first time (keyed on net, compute path, engine parameter version and cache) and replayed afterwards,
        which keeps the launch gaps off the GPU; no host synchronisation once the graph exists."""
        if self._prob is None:
            self._prob = torch.empty((self.N, self.C), dtype=torch.float32, device=self.device)
            self._value = torch.empty((self.N,), dtype=torch.float32, device=self.device)
        prob, value = self._prob, self._value
        if cache is not None and cache.net_version != net.version:
            cache.clear()                             # cached results belong to one set of weights
            cache.net_version = net.version

        def one_pass():
            if cache is not None:
                self.cached_pass(net, cache, prob, value, net_mode)
            else:
                net.forward_raw(self.planes_ptr, self.N, prob, value, net_mode)
                self.step(prob, value)

        key = (id(net), net_mode, self._version, id(cache))
        if self.use_graph and k >= 4 and self._graph_key != key:
            # one eager pass (counts), then capture the pass
            one_pass()
            k -= 1
            torch.cuda.current_stream().synchronize()
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                one_pass()
            self._graph_key = key
        if self.use_graph and self._graph_key == key:
            for _ in range(k):
                self._graph.replay()
        else:
            for _ in range(k):
                one_pass()

    # -- convenience: run until every game's budget is spent (auto_play = 0) ------
    def run_search(self, net=None, pv_fn=None, check_every: int = 16, net_mode=None, active=None, cache=None):
        """Drive step/forward until no game is busy.  ``net`` is a DeviceNet (on-device
        leaf evaluation, ``net_mode`` overrides its compute path); ``pv_fn`` is a reference-style host callable.
        ``active`` (bool / uint8 [N] device tensor, the mask given to set_roots): when fewer than ~3/4 of the slots
        search -- the arena, where each player only moves in half of the games (choose_best_player.py:48-52) -- only
        their leaves go through the network (gather planes -> forward -> scatter prob / value).  ``cache``: a
        selfplay.EvalCache shared by the games of this engine (on-device net only)."""
        assert (net is None) != (pv_fn is None)
        if self._prob is None:
            self._prob = torch.empty((self.N, self.C), dtype=torch.float32, device=self.device)
            self._value = torch.empty((self.N,), dtype=torch.float32, device=self.device)
        prob, value = self._prob, self._value
        it = 0
        self.step()
        while True:
            if pv_fn is not None:
                need = self.need_eval().cpu().numpy().astype(bool)
                if not need.any():
                    if self.busy() == 0:
                        break
                else:
                    x = self.planes().cpu().numpy().astype(np.float32)
                    p = np.zeros((self.N, self.C), np.float32)
                    v = np.zeros((self.N,), np.float32)
                    idx = np.flatnonzero(need)
                    pi, vi = pv_fn(x[idx])
                    p[idx], v[idx] = pi, vi
                    prob.copy_(torch.from_numpy(p))
                    value.copy_(torch.from_numpy(v))
            else:
                # every pass completes at least one simulation per busy game, so the largest remaining
                # budget bounds the passes still needed: run them back to back (no host sync in between),
                # then look again (games that met terminal positions finished early; `check_every` only
                # bounds the first burst when the budgets are not known to be small)
                left = max(1, int(self.sims_left().max().item()))
                it += left
                idx = None
                if active is not None:
                    idx = torch.nonzero(torch.as_tensor(active, device=self.device).reshape(-1)).reshape(-1)
                    if idx.numel() == 0 or 4 * idx.numel() > 3 * self.N:
                        idx = None
                if idx is not None:
                    planes = self.planes()
                    pc = torch.empty((idx.numel(), self.C), dtype=torch.float32, device=self.device)
                    vc = torch.empty((idx.numel(),), dtype=torch.float32, device=self.device)
                    for _ in range(left):
                        net.forward(planes.index_select(0, idx), pc, vc, net_mode)
                        prob.index_copy_(0, idx, pc)
                        value.index_copy_(0, idx, vc)
                        self.step(prob, value)
                    if self.busy() == 0:
                        break
                    continue
                self.replay_passes(left, net, cache, net_mode)
                if self.busy() == 0:
                    break
                continue
            self.step(prob, value)
        return it

    def close(self):
        self._graph = None
        self._graph_key = None
        if self.handle:
            self.lib.a5_engine_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _DevView:
    """Zero-copy view of library-owned device memory through __cuda_array_interface__."""
    _TYPES = {torch.int8: "|i1", torch.uint8: "|u1", torch.int32: "<i4", torch.float32: "<f4", torch.float64: "<f8"}

    def __init__(self, p, shape, dtype):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": self._TYPES[dtype],
                                         "data": (int(p), False), "version": 2}


def _view(p, shape, dtype, device) -> torch.Tensor:
    return torch.as_tensor(_DevView(p, shape, dtype), device=device)


def dirichlet_sample(seed: int, alpha: float, n_legal: int, n_draws: int) -> torch.Tensor:
    """n_draws Dirichlet(alpha) vectors over n_legal cells from the tree pass's sampler (a5_dirichlet_sample)."""
    eta = torch.empty((n_draws, n_legal), dtype=torch.float32, device="cuda")
    check(_lib.load().a5_dirichlet_sample(seed, alpha, n_legal, n_draws, ptr(eta), stream_ptr()))
    return eta


from .replay import parse_records  # noqa: E402,F401  (kept importable from here)
