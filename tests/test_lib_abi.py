"""CPU checks of the boundary: the C-ABI library loads here (no GPU needed to dlopen it)
and exports every symbol include/alphafive.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "alphafive.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(a5_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from alphafive_b200 import _lib, build
    build.build()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in alphafive.h but not exported"
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)
    assert lib.a5_version() == 1


def test_struct_layouts_match_header():
    from alphafive_b200 import _lib
    assert ctypes.sizeof(_lib.RecordHeader) == 32
    assert ctypes.sizeof(_lib.Config) == 104
    lib = _lib.load()
    assert lib.a5_record_stride(11) == 656 and lib.a5_record_stride(15) % 16 == 0
    names = [lib.a5_net_tensor_name(i).decode() for i in range(_lib.NUM_TENSORS)]
    from oracle import net as onet
    assert set(names) == set(onet.layer_shapes(11))
    for S in (11, 15):
        sh = onet.layer_shapes(S)
        for i, n in enumerate(names):
            size = 1
            for d in sh[n]:
                size *= d
            assert lib.a5_net_tensor_size(i, S) == size, n


def test_argument_errors_do_not_need_a_gpu():
    from alphafive_b200 import _lib
    lib = _lib.load()
    assert lib.a5_rules_terminal(None, 4, 3, 5, None, None) == -1      # S too small
    assert b"bad argument" in lib.a5_last_error()
    h = ctypes.c_void_p()
    assert lib.a5_net_create(40, 1, ctypes.byref(h)) == -1


def test_oracle_batch_terminal_matches_scalar():
    import numpy as np
    from oracle import rules
    rng = np.random.default_rng(0)
    for S in (11, 15):
        boards = np.stack([rules.random_board(rng, S) for _ in range(400)])
        want = np.array([rules.terminal_code(b) for b in boards])
        assert (rules.terminal_codes_batch(boards) == want).all()
