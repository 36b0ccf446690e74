"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/mctsb.h declares,
and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import montecarlotreesearch_b200 as m

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    m.build()
    return m.load_library()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "mctsb.h")).read()
    declared = set(re.findall(r"\b(mctsb_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(m.EXPORTS), declared ^ set(m.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_version_and_errors(lib):
    assert lib.mctsb_abi_version() == 1
    assert lib.mctsb_strerror(0) == b"ok"
    assert b"no CPU fallback" in lib.mctsb_strerror(1)


def test_wire_formats():
    # the numpy dtypes used by the tests mirror the C structs byte for byte
    assert m.QSTATE_DTYPE.itemsize == 32 and m.QSTATE_DTYPE.fields["wx"][1] == 16
    assert m.QSTATE_DTYPE.fields["move_counter"][1] == 24
    assert m.OUTCOME_DTYPE.itemsize == 16 and m.LEAFSUM_DTYPE.itemsize == 24 and m.TSTATE_DTYPE.itemsize == 4


def test_leaf_sum_conversion_is_correctly_rounded(lib):
    from fractions import Fraction
    rng = np.random.default_rng(0)
    for _ in range(200):
        s = np.zeros((), dtype=m.LEAFSUM_DTYPE)
        s["lo"] = rng.integers(0, 1 << 62, dtype=np.uint64)
        s["hi"] = rng.integers(0, 1 << 62, dtype=np.uint64)
        exact = Fraction(int(s["hi"]) * (1 << 29) + int(s["lo"]), 1 << 57)
        assert lib.mctsb_leaf_sum_to_double(s.ctypes.data) == float(exact)


def test_sm100a_cubin_only():
    out = os.popen("cuobjdump -lelf %s 2>/dev/null" % m.LIB_PATH).read()
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


@pytest.mark.skipif(m.device_count() > 0, reason="a GPU is present; the refusal path is for CPU-only hosts")
def test_no_device_means_no_compute(lib):
    h = C.c_void_p()
    assert lib.mctsb_create(C.byref(h), 0, 1) == 1      # MCTSB_ERR_NO_DEVICE
    with pytest.raises(m.BackendError) as e:
        m.RolloutBackend()
    assert e.value.status == 1


def test_tree_entry_points_refuse_to_run_without_a_backend(lib):
    """the device-resident tree hangs off a backend handle; without one (no device: no handle) every entry point is refused"""
    import numpy as np
    t = C.c_void_p()
    root = np.ascontiguousarray(m.quoridor_opening())
    assert lib.mctsb_tree_create(None, C.byref(t), root.ctypes.data, 0, 16, 4096, 1 << 20, 1.41) == 3      # MCTSB_ERR_BAD_ARG
    assert not t.value
    mv = C.c_int32(0)
    assert lib.mctsb_tree_grow(None, 16, 16, 0) == 3 and lib.mctsb_tree_best_move(None, C.byref(mv)) == 3
    assert lib.mctsb_tree_advance(None, 13) == 3 and lib.mctsb_tree_merge_root(None) == 3 and lib.mctsb_tree_set_overlap(None, 1) == 3
    assert lib.mctsb_tree_destroy(None) == 0
    assert b"device-resident tree" in lib.mctsb_strerror(9)
