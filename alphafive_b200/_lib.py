"""ctypes binding of libalphafive.so (include/alphafive.h).

The library is the product: there is no Python/CPU fallback.  ``load()`` raises if the
shared object is missing (build it with ``python -m alphafive_b200.build``) and every
wrapper raises ``A5Error`` on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libalphafive.so")
NUM_COUNTERS = 16
NUM_TENSORS = 42
NET_FP32, NET_TC, NET_SMALL = 0, 1, 2
NET_SMALL_MAX = 8


class A5Error(RuntimeError):
    pass


class Config(C.Structure):
    """a5_config (include/alphafive.h) -- field names follow the reference's config.py."""
    _fields_ = [
        ("board_size", C.c_int32), ("goal", C.c_int32), ("n_games", C.c_int32),
        ("sims", C.c_int32), ("upper_sims", C.c_int32),
        ("c_puct", C.c_float), ("dirichlet_alpha", C.c_float),
        ("init_temp", C.c_double), ("tau_decay", C.c_double), ("tau_decay_r", C.c_double),
        ("gamma", C.c_float),
        ("training", C.c_int32), ("random_a", C.c_int32), ("auto_play", C.c_int32),
        ("node_capacity", C.c_int32), ("max_inner", C.c_int32),
        ("seed", C.c_uint64), ("game_id_base", C.c_int64),
        ("record_capacity", C.c_int32), ("reserved", C.c_int32),
    ]


class RecordHeader(C.Structure):
    _fields_ = [("game_id", C.c_int64), ("game_serial", C.c_int32), ("ply", C.c_int16), ("game_len", C.c_int16),
                ("last_action", C.c_int32), ("value", C.c_float), ("weight", C.c_float), ("result", C.c_int32)]


_P = C.c_void_p
_I = C.c_int
_SIGS = {
    "a5_version": (C.c_int, []),
    "a5_last_error": (C.c_char_p, []),
    "a5_rules_terminal": (_I, [_P, _I, _I, _I, _P, _P]),
    "a5_rules_step": (_I, [_P, _P, _I, _I, _P, _P]),
    "a5_rules_legal": (_I, [_P, _I, _I, _P, _P, _P]),
    "a5_rules_inputs": (_I, [_P, _P, _I, _I, _P, _P]),
    "a5_rules_encode": (_I, [_P, _I, _I, _P, _I, _P, _P]),
    "a5_rules_decode": (_I, [_P, _I, _I, _I, _P, _P]),
    "a5_engine_create": (_I, [C.POINTER(Config), C.POINTER(_P)]),
    "a5_engine_destroy": (_I, [_P]),
    "a5_engine_reset": (_I, [_P, _P]),
    "a5_engine_set_roots": (_I, [_P, _P, _P, _P, _P, _P]),
    "a5_engine_step": (_I, [_P, _P, _P, _P]),
    "a5_engine_set_mode": (_I, [_P, _I, _I]),
    "a5_engine_set_budget": (_I, [_P, _I, _I]),
    "a5_engine_planes": (_P, [_P]),
    "a5_engine_need_eval": (_P, [_P]),
    "a5_engine_sims_left": (_P, [_P]),
    "a5_engine_busy": (_I, [_P, C.POINTER(C.c_int32), _P]),
    "a5_engine_finish_move": (_I, [_P, _P, _P, _P]),
    "a5_engine_collect_moves": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P]),
    "a5_engine_submit_roots": (_I, [_P, _I, _P, _P, _P, _P, _P]),
    "a5_engine_root_stats": (_I, [_P, _P, _P, _P, _P, _P]),
    "a5_engine_node_stats": (_I, [_P, _P, _P, _P, _P, _P, _P]),
    "a5_engine_tau": (_P, [_P]),
    "a5_engine_get_roots": (_I, [_P, _P, _P, _P]),
    "a5_dirichlet_sample": (_I, [C.c_uint64, C.c_float, _I, _I, _P, _P]),
    "a5_engine_table_dump": (_I, [_P, _I, _P, _P, _I, C.POINTER(C.c_int32), _P]),
    "a5_record_stride": (_I, [_I]),
    "a5_net_forward_parts": (_I, [_P, _P, _I, _P, _P, _I, _P]),
    "a5_net_set_sm_limit": (_I, [_P, _I, _I]),
    "a5_replay_sample": (_I, [_P, _I, _P, _P, _P, _I, _P, _P, _P, _P, _P]),
    "a5_engine_harvest": (_I, [_P, _P, _I, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P]),
    "a5_engine_counters": (_I, [_P, C.POINTER(C.c_int64), _P]),
    "a5_net_tensor_name": (C.c_char_p, [_I]),
    "a5_net_tensor_size": (C.c_int64, [_I, _I]),
    "a5_net_create": (_I, [_I, _I, C.POINTER(_P)]),
    "a5_net_destroy": (_I, [_P]),
    "a5_net_set_weights": (_I, [_P, C.POINTER(_P), _P]),
    "a5_net_forward": (_I, [_P, _P, _I, _P, _P, _I, _P]),
    "a5_engine_step_served": (_I, [_P, _P, _P, _P, _P]),
    "a5_evalcache_create": (_I, [_I, _I, _I, _I, C.POINTER(_P)]),
    "a5_evalcache_destroy": (_I, [_P]),
    "a5_evalcache_clear": (_I, [_P, _P]),
    "a5_evalcache_lookup": (_I, [_P, _P, _P, _P, _P, _P, _P]),
    "a5_evalcache_planes": (_P, [_P]),
    "a5_evalcache_prob": (_P, [_P]),
    "a5_evalcache_value": (_P, [_P]),
    "a5_evalcache_commit": (_I, [_P, _P, _P, _P]),
    "a5_evalcache_stats": (_I, [_P, C.POINTER(C.c_int64), _P]),
}

EXPORTS = tuple(_SIGS)
_lib = None


def load():
    """Load the library (once).  Raises A5Error when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise A5Error(f"{LIB_PATH} is missing: run `python -m alphafive_b200.build` "
                      "(there is no CPU fallback for the self-play hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        msg = load().a5_last_error().decode(errors="replace")
        raise A5Error(f"libalphafive status {status}: {msg}")


def ptr(t):
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
