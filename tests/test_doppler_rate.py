"""Doppler-RATE search dimension.  The reference prepares `complexHeterodyne`
(cuda_kernels.cu:755-778, demodulator_base.py:388) and never calls it; `pcs_heterodyne` is the same fp32 statement applied to
the uploaded chunk, and a rate search is one ordinary Doppler search per hypothesis on the de-chirped chunk.

CPU: the oracle's restatement against the closed form.  GPU: de-chirped chunk <= 1e-5, energies per hypothesis <= 1e-4,
the same winning hypothesis, and -- at the winning hypothesis -- identical shift / timing bin / bits as the oracle."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import signals as S
from tests.helpers import RADIO, conf_variant, protocol_for, rel_err

FS = 153600.0


@pytest.mark.parametrize("rate", [50.0, -200.0, 30000.0])
def test_oracle_heterodyne_is_the_fp32_chirp(rate):
    rng = np.random.RandomState(1)
    N = 2 ** 15
    x = (rng.randn(N) + 1j * rng.randn(N)).astype(np.complex64)
    a = O.rate_to_a(rate, FS)
    y = O.heterodyne(x, a)
    n = np.arange(N, dtype=np.float64)
    ref = x.astype(np.complex128) * np.exp(1j * float(a) * n * n)
    # fp32 phase: |theta| reaches a N^2, so the error is ~ ulp(theta_max) in radians
    tol = 2e-6 + 2 * np.spacing(np.float32(abs(float(a)) * N * N))
    assert y.dtype == np.complex64 and rel_err(y, ref) < tol
    # de-chirping a chirp gives back the original
    chirped = S.doppler_rate(x, rate, FS).astype(np.complex64)
    assert rel_err(O.heterodyne(chirped, a), x) < tol + 1e-6


def _chirped_chunk(rate, seed=3, snr=15):
    conf = conf_variant("benchmark/bench_GMSK.json", blockSize=15)
    sig, _ = S.get_padded_packet("GMSK", 16, FS, pad=100)
    N = 2 ** 15
    rng = np.random.RandomState(seed)
    x = S.doppler_rate(sig[:N], rate, FS)
    x = x + 10 ** (-snr / 20) * (rng.randn(N) + 1j * rng.randn(N)) / np.sqrt(2)
    return conf, x.astype(np.complex64)


@pytest.mark.gpu
@pytest.mark.parametrize("rate", [30000.0, -45000.0])
def test_rate_search_matches_oracle(rate):
    from pycusdr_b200.demodulator import UHF
    conf, x = _chirped_chunk(rate)
    P = protocol_for(conf)
    dem, orc = UHF.Demodulator(conf, P, RADIO), O.OracleDemodulator(conf, P, RADIO)
    rates = [-45000.0, -30000.0, -15000.0, 0.0, 15000.0, 30000.0, 45000.0]
    raw = dem.get_signalBufferHostPointer()
    # 1. the de-chirped chunk itself
    raw[:] = x
    dem.uploadToGPU(raw)
    a = O.rate_to_a(rate, FS)
    dem._engine.heterodyne(a)
    assert rel_err(dem._engine.chunk(), O.heterodyne(x, a)) < 1e-5
    dem._engine.heterodyne(0.0)
    np.testing.assert_array_equal(dem._engine.chunk(), x)                 # a = b = c = 0 restores the uploaded chunk
    # 2. the search over the hypotheses
    out = dem.findUHFRates(raw, rates)
    best_o, metrics_o, Es = O.search_rates(x, orc.masks, orc.doppCyperSymNorm, orc.SUM_ALL_MASKS_PYTHON, rates, FS)
    assert out[0] == rates[best_o] == rate
    np.testing.assert_allclose(out[1], metrics_o, rtol=1e-4)
    assert rel_err(dem.last["E"], Es[best_o]) < 1e-4
    # the matched filters are short (48 taps: ~3 kHz wide), so the metric is a shallow function of the rate: the true
    # hypothesis beats "no de-chirp" by 14 % (30 kHz/s) / 36 % (45 kHz/s) and its neighbours by ~4 %
    assert out[1][rates.index(rate)] > 1.1 * out[1][rates.index(0.0)]
    # 3. demodulation at the winning hypothesis: the oracle on the oracle's de-chirped chunk
    bits = dem.demodulate()[0]
    xo = O.heterodyne(x, a)
    orc.get_signalBufferHostPointer()[:] = xo
    orc.uploadAndFindCarrier(orc.get_signalBufferHostPointer())
    bo = orc.demodulate()[0]
    assert dem.last["shift"] == orc.last["shift"] and dem.last["timing"][0] == orc.last["timing"][0]
    live = orc.last["mag"] > 1e-6 * orc.last["mag"].max()
    n_diff = int(np.sum(dem.last["sym"][live] != orc.last["sym"][live]))
    assert n_diff <= 2e-3 * live.sum()
    if n_diff == 0:
        np.testing.assert_array_equal(bits, bo)
