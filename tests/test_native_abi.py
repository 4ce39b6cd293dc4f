"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/pycusdr_b200.h
declares, the ctypes binding covers the same list, and the product path fails loudly without a GPU (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from pycusdr_b200 import _native
from tests.helpers import RADIO, have_gpu, load_conf, protocol_for

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pycusdr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcs_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_the_binding_binds():
    assert _declared() == sorted(_native.SYMBOLS)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_native.LIB_PATH), "build with `make -C pycusdr_b200/csrc`"
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} not exported"
    assert _native.load().pcs_abi_version() == _native.ABI_VERSION


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(_native.Config) == 16 * 4
    assert ctypes.sizeof(_native.Result) == 96
    assert ctypes.sizeof(_native.PlanInfo) == 10 * 4 + 8


def test_product_path_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "pycusdr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)


@pytest.mark.skipif(have_gpu(), reason="this checks the behaviour on a machine WITHOUT a GPU")
def test_no_gpu_means_an_exception_not_a_fallback():
    from pycusdr_b200.demodulator import UHF
    conf = load_conf("benchmark/bench_GMSK.json")
    with pytest.raises(_native.NativeError) as e:
        UHF.Demodulator(conf, protocol_for(conf), RADIO)
    assert e.value.code in (-3, -2)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_bad_arguments_are_rejected_before_touching_the_device():
    lib = _native.load()
    assert lib.pcs_create(None, None, None, None) == -1
    assert b"null" in lib.pcs_last_error()
    cfg = _native.Config(abi_version=99)
    h = ctypes.c_void_p()
    shifts = np.zeros(1, np.int32)
    masks = np.zeros((1, 4096), np.complex64)
    rc = lib.pcs_create(ctypes.byref(cfg), shifts.ctypes.data_as(ctypes.c_void_p), masks.ctypes.data_as(ctypes.c_void_p),
                        ctypes.byref(h))
    assert rc == -1 and b"ABI version" in lib.pcs_last_error()
    assert lib.pcs_destroy(None) == 0


def test_window_mean_of_magnitudes_matches_numpy():
    """pcs_mean_abs_c64 (the computeSNR window means, dem_base:657-661) against np.mean(np.abs(.)): equal to float32
    rounding (NumPy's own result moves in the last ulp with the host's SIMD dispatch)."""
    rng = np.random.RandomState(4)
    for n in (1, 2, 7, 10, 170, 1000, 8192):
        z = ((rng.randn(n) + 1j * rng.randn(n)) * 10.0 ** rng.uniform(-3, 3)).astype(np.complex64)
        got = _native.mean_abs_c64(z)
        assert got.dtype == np.float32
        np.testing.assert_allclose(got, np.mean(np.abs(z.astype(np.complex128))), rtol=2e-7)
        np.testing.assert_allclose(got, np.mean(np.abs(z)), rtol=1e-6)
    with pytest.raises(_native.NativeError):
        _native.mean_abs_c64(np.zeros(0, np.complex64))


def test_gap_filling_of_clipped_indices_matches_the_numpy_statement():
    """pcs_fill_gaps against the reference's own gap-filling statements (dem_base:686-705) on random index sets."""
    rng = np.random.RandomState(11)
    N = 4096
    for trial in range(200):
        n = int(rng.randint(0, 60))
        over = np.unique(rng.randint(0, N, size=n)) if trial % 3 else np.unique(np.concatenate(
            [rng.randint(0, N, size=max(n // 4, 1)) + k for k in (0, 1, 3, 50, 99, 100, 101)]) % N)
        got = _native.fill_gaps(over, 100, N)
        if len(over) == 0:
            assert len(got) == 0
            continue
        # the reference's statements, verbatim in meaning
        diffPeaks = np.diff(over)
        gapsAll = np.where(diffPeaks > 1)[0]
        gaps = np.where(diffPeaks[gapsAll] < 100)[0]
        gapsLen, gapsIdx = diffPeaks[gapsAll[gaps]], gapsAll[gaps]
        pp = np.zeros(N, dtype=np.int8)
        pp[over] = 1
        for i in range(len(gapsLen)):
            pp[over[gapsIdx[i]]:over[gapsIdx[i]] + gapsLen[i]] = 1
        np.testing.assert_array_equal(got, np.where(pp == 1)[0])
    with pytest.raises(_native.NativeError):
        _native.fill_gaps(np.array([5, 5], np.int64), 100, N)
