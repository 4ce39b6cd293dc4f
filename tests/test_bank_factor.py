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


def _incomplete_and_other_banks():
    from pycusdr_b200.protocol.FSK2_base import fsk_phase_templates
    N, sps = 2 ** 14, 128

    def spectra(tmpl):
        m = np.zeros((len(tmpl), N), dtype=np.complex128)
        for i, tp in enumerate(tmpl):
            m[i, :len(tp)] = tp
        return np.conj(np.fft.fft(m, axis=1)).astype(np.complex64)
    full = [np.array([(k >> 2) & 1, (k >> 1) & 1, k & 1]) for k in range(8)]
    n = np.arange(sps)
    tones = [np.exp(2j * np.pi * f * n / sps) for f in (-1.0, 0.5, 1.5)]
    return N, {
        "incomplete": spectra(fsk_phase_templates([p for k, p in enumerate(full) if k != 5], sps, 0.5)),
        "three_tones": spectra([np.concatenate((tones[a], np.exp(0.3j * (a + 2 * b)) * tones[b])) for a in range(3) for b in range(3)]),
        # complete, but the first segment's constant depends on the LAST symbol: no shared partial sums
        "no_prefix": spectra([np.concatenate((np.exp(0.4j * b) * tones[a], tones[b])) for a in (0, 1) for b in (0, 1)]),
    }


def test_code_order_of_the_cc11xx_bank_and_shared_prefix_sums():
    """pcs_bank_code_order: CC11xx is a complete binary bank whose coefficients depend on the code's prefix only (form 3).  The
    combination the kernel then runs -- level j adds c[prefix][j] u_{bit j}[i + j S] to the partial sum of the prefix one bit
    shorter -- must give the same y as the plain J-term sum in mask order."""
    orc, masks, shifts = _bank("CC11xx.json", 14, 5)
    lp, ln = _support(masks, masks.shape[1])
    f = _native.factorise_bank(masks, lp, ln, shifts, 11)
    form, code_mask, coef = _native.bank_code_order(f)
    assert form == 3 and sorted(code_mask.tolist()) == list(range(8))
    J, S = f["J"], f["S"]
    for code, m in enumerate(code_mask):
        assert [int(f["sel"][m, j]) for j in range(J)] == [(code >> j) & 1 for j in range(J)]
        np.testing.assert_array_equal(coef[:, code, :], f["coef"][:, m, :])
    form2, _, _ = _native.bank_code_order(f, allow_shared_sums=False)
    assert form2 == 2
    # the kernel's tree evaluation against the plain sum, on random u
    rng = np.random.RandomState(3)
    u = rng.randn(2, 1024) + 1j * rng.randn(2, 1024)
    i = np.arange(512)
    d = 2
    for code, m in enumerate(code_mask):
        plain = sum(f["coef"][d, m, j].astype(np.complex128) * u[f["sel"][m, j], i + j * S] for j in range(J))
        part = 0
        for j in range(J):
            pre = code & ((2 << j) - 1)
            part = part + coef[d, pre, j].astype(np.complex128) * u[(pre >> j) & 1, i + j * S]
        assert np.max(np.abs(part - plain)) <= 1e-6 * np.max(np.abs(plain))


def test_code_order_leaves_other_banks_in_the_general_form():
    N, banks = _incomplete_and_other_banks()
    shifts = (((np.arange(6) - 3) * 37) % N).astype(np.int32)
    want = {"incomplete": 1, "three_tones": 1, "no_prefix": 2}
    for name, masks in banks.items():
        lp, ln = _support(masks, N)
        f = _native.factorise_bank(masks, lp, ln, shifts, 11)
        assert f is not None, name
        form, sel, coef = _native.bank_code_order(f)
        assert form == want[name], name
        if form == 1:
            np.testing.assert_array_equal(sel, f["sel"])
            np.testing.assert_array_equal(coef, f["coef"])
