"""Segment factorisation of the long-filter bank (pcs_factorise_bank, pycusdr_b200/csrc/bank_factor.cu; CPU only).

The tables the library hands search_fb_kernel are checked here in float64 against the oracle's literal statement of the
reference (kern:339-373 multiply, inverse FFT, kern:421-480 |.|^2) on the reference's own filter banks
(pyCuSDR/protocol/FSK2_base.py:17-46 for CC11xx): a NumPy emulation of exactly what the kernel does -- R inverse B-point
transforms of (block spectrum x basis spectrum), then y_m[i] = sum_j c[d][m][j] u_{sel[m][j]}[i + j S] on the block's
valid outputs -- must reproduce the Nfft-point circular correlation on every sample.
"""
import numpy as np
import pytest

from oracle import oracle as O
from pycusdr_b200 import _native
from tests.helpers import RADIO, conf_variant, protocol_for
from tests.test_restructuring_identities import _support


def _bank(cfg, blockSize, bins):
    conf = conf_variant(cfg, blockSize=blockSize, doppCarrierSteps=bins)
    orc = O.OracleDemodulator(conf, protocol_for(conf), RADIO)
    masks = np.ascontiguousarray(orc.masks, dtype=np.complex64)
    shifts = np.asarray(orc.doppCyperSymNorm, dtype=np.int64) % (2 ** blockSize)
    return orc, masks, shifts.astype(np.int32)


def test_cc11xx_bank_is_two_tones_times_constants():
    orc, masks, shifts = _bank("CC11xx.json", 14, 9)
    N = masks.shape[1]
    lp, ln = _support(masks, N)
    assert lp + ln + 1 == 384                                     # 3 symbols x 128 samples (FSK2_base.py:17-46)
    f = _native.factorise_bank(masks, lp, ln, shifts, 11)
    assert f is not None and (f["S"], f["J"], f["R"]) == (128, 3, 2)
    assert f["sel"].shape == (8, 3) and set(np.unique(f["sel"])) == {0, 1}
    # every template is one of the 2^3 tone sequences: the selectors enumerate all of them
    assert len({tuple(r) for r in f["sel"]}) == 8
    # unit-modulus templates: |c| is the same for every segment
    c0 = np.abs(f["coef"][0])
    np.testing.assert_allclose(c0, c0.flat[0], rtol=1e-5)


@pytest.mark.parametrize("cfg,blockSize,logB", [("CC11xx.json", 14, 11), ("CC11xx.json", 13, 10), ("CC11xx.json", 14, 12)])
def test_factorised_search_equals_the_full_transform(cfg, blockSize, logB):
    orc, masks, shifts = _bank(cfg, blockSize, 5)
    M, N = masks.shape
    B = 1 << logB
    lp, ln = _support(masks, N)
    L = lp + ln + 1
    f = _native.factorise_bank(masks, lp, ln, shifts, logB)
    assert f is not None
    S, J, R, sel = f["S"], f["J"], f["R"], f["sel"]
    rng = np.random.RandomState(blockSize + logB)
    x = (rng.randn(N) + 1j * rng.randn(N)) * 0.5
    X = np.fft.fft(x)
    k = np.arange(N)
    V = B - L + 1
    nblk = (N + V - 1) // V
    Mk = masks.astype(np.complex128)
    for d, s in enumerate(shifts):
        y_ref = np.fft.ifft(X[None, :] * Mk[:, (k - s) % N], axis=1) * N      # shift charged to the filter (same |y|)
        y = np.zeros((M, N), dtype=np.complex128)
        for blk in range(nblk):
            n0 = blk * V
            idx = (n0 - lp + np.arange(B)) % N
            Xb = np.fft.fft(x[idx])                                           # block_spectra_kernel
            u = np.fft.ifft(Xb[None, :] * f["basis_spec"][d].astype(np.complex128), axis=1) * B    # R transforms
            vlen = min(V, N - n0)
            i = lp + np.arange(vlen)                                          # valid outputs of the block
            for m in range(M):
                acc = np.zeros(vlen, dtype=np.complex128)
                for j in range(J):
                    assert (i + j * S).max() < B
                    acc += f["coef"][d, m, j].astype(np.complex128) * u[sel[m, j], i + j * S]
                y[m, n0:n0 + vlen] = acc
        scale = np.max(np.abs(y_ref))
        assert np.max(np.abs(y - y_ref)) <= 3e-6 * scale
        np.testing.assert_allclose(np.sum(np.abs(y) ** 2, axis=1), np.sum(np.abs(y_ref) ** 2, axis=1), rtol=1e-5)


@pytest.mark.parametrize("cfg", ["benchmark/bench_GMSK.json", "benchmark/bench_BPSK.json"])
def test_banks_without_segment_structure_are_left_alone(cfg):
    """Gaussian-filtered / root-raised-cosine templates are not piecewise multiples of a few segments."""
    orc, masks, shifts = _bank(cfg, 12, 3)
    lp, ln = _support(masks, masks.shape[1])
    assert _native.factorise_bank(masks, lp, ln, shifts, 10) is None


def test_bad_geometry_is_rejected():
    masks = np.zeros((2, 1024), dtype=np.complex64)
    with pytest.raises(_native.NativeError):
        _native.factorise_bank(masks, 600, 600, np.zeros(2, np.int32), 10)       # support longer than the block
