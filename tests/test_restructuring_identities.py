"""The algebra behind the CUDA search kernels, checked in float64 against the oracle's literal
restatement of the reference (kern:339-373 multiply, inverse FFT, kern:421-480 abs-sum) on the reference's own filter
banks.  CPU only; the GPU parity tests check the kernels, these check that what the kernels are asked to compute is the
same function of the input as what the reference computes.

  1. the spectrum shift can be charged to the filter:  |IFFT(X[(k+s)%N] Mk[k])|^2 == |IFFT(X[k] Mk[(k-s)%N])|^2;
  2. overlap-save with B-point transforms and the B-point filter spectra Mk[(k N/B - s) % N] N/B reproduces the
     Nfft-point circular correlation on every sample, for the 256-point plan (short filters) and the 2048-point plan
     (CC11xx, 384 taps);
  3. Parseval: sum_n |y|^2 == N sum_k |X[(k+s)%N]|^2 |Mk[k]|^2 (the labelled variant).
"""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import RADIO, conf_variant, protocol_for


def _setup(cfg, blockSize, bins=5):
    conf = conf_variant(cfg, blockSize=blockSize, doppCarrierSteps=bins)
    orc = O.OracleDemodulator(conf, protocol_for(conf), RADIO)
    N = 2 ** blockSize
    rng = np.random.RandomState(blockSize)
    x = (rng.randn(N) + 1j * rng.randn(N)) * 0.5
    return orc, N, x, np.fft.fft(x)


def _support(masks, N):
    """Taps of g_m = IFFT(Mk_m) above 1e-5 of the peak amplitude: n = 0..Lpos and n = -Lneg..-1 (what pcs_create measures)."""
    lp = ln = 0
    for m in range(masks.shape[0]):
        g = np.abs(np.fft.ifft(masks[m].astype(np.complex128)))
        idx = np.flatnonzero(g > 1e-5 * g.max())
        lp = max(lp, int(idx[idx < N // 2].max(initial=0)))
        neg = idx[idx >= N // 2]
        if neg.size:
            ln = max(ln, int(N - neg.min()))
    return lp, ln


@pytest.mark.parametrize("cfg,blockSize", [("benchmark/bench_GMSK.json", 13), ("benchmark/bench_FSK.json", 12), ("CC11xx.json", 14)])
def test_shift_can_be_charged_to_the_filter(cfg, blockSize):
    orc, N, x, X = _setup(cfg, blockSize)
    Mk = orc.masks.astype(np.complex128)
    k = np.arange(N)
    for s in orc.doppCyperSymNorm:
        y_ref = np.fft.ifft(X[(k + s) % N][None, :] * Mk, axis=1) * N            # the reference's statement
        y_flt = np.fft.ifft(X[None, :] * Mk[:, (k - s) % N], axis=1) * N         # shift moved onto the filter
        np.testing.assert_allclose(np.abs(y_flt) ** 2, np.abs(y_ref) ** 2, rtol=1e-9, atol=1e-9 * np.max(np.abs(y_ref)) ** 2)
        # and they differ exactly by the unit-modulus rotation exp(+2 pi i s n / N)
        np.testing.assert_allclose(y_flt, y_ref * np.exp(2j * np.pi * s * k / N)[None, :], rtol=1e-8,
                                   atol=1e-9 * np.max(np.abs(y_ref)))


@pytest.mark.parametrize("cfg,blockSize,B", [("benchmark/bench_GMSK.json", 13, 256), ("benchmark/bench_BPSK.json", 12, 256),
                                             ("CC11xx.json", 14, 2048), ("benchmark/bench_GMSK.json", 13, 512)])
def test_overlap_save_with_shifted_filter_spectra_equals_the_full_transform(cfg, blockSize, B):
    orc, N, x, X = _setup(cfg, blockSize, bins=3)
    Mk = orc.masks.astype(np.complex128)
    M = Mk.shape[0]
    lp, ln = _support(orc.masks, N)
    L = lp + ln + 1
    V = B - L + 1
    assert V >= B // 2, "plan rule of pcs_create"
    dec = N // B
    k = np.arange(N)
    kb = np.arange(B)
    nblk = -(-N // V)
    for s in orc.doppCyperSymNorm:
        y_ref = np.fft.ifft(X[(k + s) % N][None, :] * Mk, axis=1) * N
        G = Mk[:, (kb * dec - s) % N] * dec                                     # shifted_filters kernels
        got = np.zeros((M, N))
        for b in range(nblk):
            n0 = b * V
            xb = np.fft.fft(x[(n0 - lp + kb) % N])                              # block_spectra kernels
            yb = np.fft.ifft(xb[None, :] * G, axis=1) * B
            vlen = min(V, N - n0)
            got[:, n0:n0 + vlen] = np.abs(yb[:, lp:lp + vlen]) ** 2
        ref = np.abs(y_ref) ** 2
        # the filters are measured to be zero outside [-ln, lp] at 1e-5 of the peak tap: that is the only approximation
        assert np.max(np.abs(got - ref)) <= 2e-5 * np.max(ref)
        np.testing.assert_allclose(got.sum(axis=1), ref.sum(axis=1), rtol=1e-5)
        assert np.array_equal(np.argmax(got, axis=1), np.argmax(ref, axis=1))


@pytest.mark.parametrize("cfg,blockSize", [("benchmark/bench_GMSK.json", 12), ("CC11xx.json", 13)])
def test_parseval_form_of_the_energy(cfg, blockSize):
    orc, N, x, X = _setup(cfg, blockSize)
    Mk = orc.masks.astype(np.complex128)
    k = np.arange(N)
    for s in orc.doppCyperSymNorm:
        y = np.fft.ifft(X[(k + s) % N][None, :] * Mk, axis=1) * N
        lhs = (np.abs(y) ** 2).sum(axis=1)
        rhs = N * (np.abs(X[(k + s) % N]) ** 2 * np.abs(Mk) ** 2).sum(axis=1)
        np.testing.assert_allclose(lhs, rhs, rtol=1e-10)
