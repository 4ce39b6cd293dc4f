"""Pins the oracle's restatement of the per-chunk device kernels to the reference's OWN device code.

``tests/golden/refgpu_*.npz`` were produced on a B200 by ``oracle/ref_gpu/make_golden_gpu.py``: the unmodified
``pyCuSDR/demodulator/cuda_kernels.cu`` (compiled for sm_100a) + cuFFT, launched with the reference's call sequence
and shapes.  Here the same seeded inputs go through ``oracle/oracle.py`` on the CPU.

Tolerances: anything that crosses an FFT is compared at fp32-FFT accuracy (cuFFT and pocketfft round differently):
energies 1e-4 relative, Doppler estimate 2e-3 bins, timing phase 1e-4 rad; integer decisions (selected shift, timing
bin) must be identical; symbol decisions / bits identical except where two candidates tie to within FFT rounding.
Kernels without an FFT in them (findDopplerEst, sumXCorrBuffMasks, findCodeRateAndPhase, findCentres) are checked
on the reference's own intermediate buffers and must match bit for bit.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle.ref_gpu import make_golden_gpu as G
from tests.helpers import RADIO, rel_err


@pytest.fixture(scope="module")
def streams(golden_dir):
    path = os.path.join(golden_dir, "refgpu_streams.npz")
    if not os.path.exists(path):
        pytest.skip("refgpu_streams.npz not generated yet")
    return np.load(path)


@pytest.fixture(scope="module")
def small(golden_dir):
    path = os.path.join(golden_dir, "refgpu_small.npz")
    if not os.path.exists(path):
        pytest.skip("refgpu_small.npz not generated yet")
    return np.load(path)


def _chunks_of(streams, case):
    return sorted({int(k.split("/")[1]) for k in streams.files if k.startswith(case + "/")})


@pytest.mark.parametrize("case", list(G.STREAM_CASES))
def test_stream_matches_reference_kernels(streams, case):
    conf, P, sig, keep, backend = G.stream_for(case)
    orc = O.OracleDemodulator(conf, P, RADIO, backend=backend)
    N, ovl = orc.Nfft, orc.sigOverlap
    step = N - ovl
    raw = orc.get_signalBufferHostPointer()
    raw[:] = 0
    want = _chunks_of(streams, case)
    assert len(want) >= 4
    sym_total = sym_bad = bit_total = bit_bad = 0
    for c in range(max(want) + 1):
        raw[ovl:] = sig[c * step:(c + 1) * step]
        f = orc.uploadAndFindCarrier(raw)
        b = orc.demodulate()
        if c in want:
            g = lambda name: streams[f"{case}/{c}/{name}"]          # noqa: E731
            L = orc.last
            assert abs(np.sum(raw.astype(np.complex128)) - g("xsum")[0]) < 1e-6 * N, "input stream differs"
            if backend == "UHF":
                assert rel_err(L["E"], g("E")) < 1e-4
                assert abs(float(L["res"][0]) - float(g("res")[0])) < 2e-3
                assert abs(float(L["res"][1]) - float(g("res")[1])) < 2e-3
                assert f[0] == pytest.approx(g("ret")[0], abs=0.5)        # frequency offset, Hz
                np.testing.assert_allclose(f[3], g("ret")[2], rtol=1e-4, atol=1e-4, equal_nan=True)   # SNR
            assert L["shift"] == int(g("shift")[0])
            assert L["timing"][0] == g("timing")[0]
            dphi = np.angle(np.exp(1j * (float(L["timing"][1]) - float(g("timing")[1]))))
            assert abs(dphi) < 1e-4
            rs, rc, rm = g("sym").astype(np.int32), g("centres"), g("mag")
            assert len(rs) == len(L["sym"])
            sym_total += len(rs)
            bad = (rs != L["sym"]) | (rc != L["centres"])
            sym_bad += int(bad.sum())
            ok = ~bad
            assert rel_err(L["mag"][ok], rm[ok]) < 1e-4
            rb = np.unpackbits(g("bits"))[:int(g("nbits")[0])]
            if len(rb) == len(b[0]):
                bit_total += len(rb)
                bit_bad += int(np.sum(rb != b[0]))
            else:
                bit_bad += abs(len(rb) - len(b[0]))
                bit_total += len(rb)
            np.testing.assert_array_equal(np.asarray(orc.clippedPeakIPure, dtype=np.int64), g("clipped"))
        raw[:ovl] = raw[-ovl:]
    # near-ties resolved by FFT rounding are the only admissible differences
    assert sym_bad <= 2e-3 * sym_total, f"{sym_bad} of {sym_total} symbol decisions differ"
    assert bit_bad <= 2e-3 * bit_total, f"{bit_bad} of {bit_total} output bits differ"


def test_c1_plumbing_chunk(streams):
    from tests.helpers import conf_variant, protocol_for
    conf = conf_variant("CC11xx.json")
    P = protocol_for(conf)
    orc = O.OracleDemodulator(conf, P, RADIO)
    x = G.c1_chunk(conf)
    g = lambda name: streams[f"c1/0/{name}"]                          # noqa: E731
    assert abs(np.sum(x.astype(np.complex128)) - g("xsum")[0]) < 1e-6 * len(x)
    orc.get_signalBufferHostPointer()[:] = x
    f = orc.uploadAndFindCarrier(orc.get_signalBufferHostPointer())
    b = orc.demodulate()
    L = orc.last
    assert rel_err(L["E"], g("E")) < 1e-4
    assert L["shift"] == int(g("shift")[0])
    assert L["timing"][0] == g("timing")[0]
    assert f[0] == pytest.approx(g("ret")[0], abs=0.5)
    rs = g("sym").astype(np.int32)
    live = g("mag") > 1e-6 * g("mag").max()          # the packet occupies part of the chunk; the rest is noise
    assert np.mean(rs[live] != L["sym"][live]) < 5e-3


# ---- kernel-by-kernel checks on the reference's own intermediate buffers ------------------------------------
def test_forward_fft_and_surface(small):
    X = O.forward_fft(small["x"])
    assert rel_err(X, small["X"]) < 2e-6
    from tests.helpers import conf_variant, protocol_for
    conf = conf_variant("benchmark/bench_GMSK.json", blockSize=12, doppCarrierSteps=6)
    masks = protocol_for(conf).get_filter(4096, 16, 3)[1]
    for d, s in enumerate(small["shifts"]):
        y = O.surface_rows(small["X"], masks, s)
        assert np.max(np.abs(y - small["surface"][d])) < 1e-5 * np.max(np.abs(y))


def test_energy_reduction_on_reference_surface(small):
    surf = small["surface"].astype(np.complex64)
    E = np.sum(O.abs2(surf).astype(np.float64) / 262144.0, axis=2)
    assert rel_err(E, small["E"]) < 2e-6          # float atomics / tree order only


def test_find_doppler_est_bit_exact(small, streams):
    r = O.find_doppler_est(small["E"], 6, 0, False)
    assert r.tobytes() == small["res"].tobytes()
    n = 0
    for k in streams.files:
        if k.endswith("/E"):
            case = k.split("/")[0]
            if case == "c1":
                D, off, sum_all = 64, 0, True
            else:
                over = G.STREAM_CASES[case][6]
                off = 1 if "noise_measure_offset_Hz" in over else 0
                D, sum_all = 64, over.get("sum_all", True)
            r = O.find_doppler_est(streams[k], D, off, sum_all)
            ref = streams[k[:-1] + "res"]
            assert r[0].tobytes() == ref[0].tobytes(), k
            assert abs(float(r[1]) - float(ref[1])) <= 2e-6 * max(1.0, abs(float(ref[1]))), k   # log10f: 1 ulp
            n += 1
    assert n > 30


def test_sum_masks_bit_exact(small):
    y = small["y"]
    p = np.zeros(y.shape[1], dtype=np.float32)
    for m in range(y.shape[0]):
        p = (p + O.abs2_fma(y[m], ORDER)).astype(np.float32)
    assert p.tobytes() == small["p"].tobytes()


def test_timing_kernel_on_reference_spectrum(small):
    N = len(small["x"])
    Pf = np.fft.rfft(small["p"].astype(np.float64))
    assert rel_err(Pf, small["Pf"]) < 2e-6
    iH, iL = O.code_rate_band(N, 16)
    r = O.find_code_rate_and_phase(small["Pf"], iH, iL - iH, abs2=lambda z: O.abs2_fma(z, ORDER))
    assert r[0] == small["timing"][0]
    assert r[2].tobytes() == small["timing"][2].tobytes()
    assert abs(float(r[1]) - float(small["timing"][1])) < 5e-7        # atan2f vs NumPy arctan2: <= 2 ulp


def test_find_centres_bit_exact(small):
    N = len(small["x"])
    spSym, off = O.code_rate_host(small["timing"], N)
    sym, centres, mag = O.find_centres(O.abs2_fma(small["y"], ORDER), spSym, off, N, 7, 8)
    np.testing.assert_array_equal(sym, small["sym"])
    np.testing.assert_array_equal(centres, small["centres"])
    assert mag.tobytes() == small["mag"].tobytes()


ORDER = "xy"
