"""The CUDA path against the reference's OWN device code on the same GPU: pyCuSDR/demodulator/cuda_kernels.cu compiled
unmodified for sm_100a (oracle/_ref/*.cubin, built by oracle/ref_gpu/Makefile where /root/reference is mounted) + cuFFT,
launched with the reference's call sequence (oracle/ref_gpu/driver.py).  Same tolerances as against the oracle:
energies 1e-4 relative (the reference's float atomics are order dependent), identical shift and timing bin, symbol
decisions identical except where two candidates tie to within fp32-FFT rounding, identical output bits."""
import numpy as np
import pytest

from oracle import signals as S
from oracle import oracle as O
from oracle.ref_gpu import driver as R
from tests.helpers import RADIO, conf_variant, load_conf, protocol_for, rel_err

pytestmark = pytest.mark.gpu


def _need(M, W, sum_all):
    if not R.available(M, W, sum_all, 0):
        pytest.skip("oracle/_ref cubin for this variant is not built (needs /root/reference at build time)")


@pytest.mark.parametrize("mod,cfg,snr,sum_all", [("GMSK", "benchmark/bench_GMSK.json", 12, True),
                                                  ("FSK", "benchmark/bench_FSK.json", 9, True),
                                                  ("GMSK", "benchmark/bench_GMSK.json", 15, False),
                                                  ("BPSK", "benchmark/bench_BPSK.json", 10, True)])
def test_stream_against_reference_kernels(mod, cfg, snr, sum_all):
    from pycusdr_b200.demodulator import UHF
    conf = load_conf(cfg)
    P = protocol_for(conf)
    P.SUM_ALL_MASKS_PYTHON = sum_all
    _need(P.get_filter(4096, 16, conf["GPU"]["UHF"]["xcorrMaskSize"])[0], 7, sum_all)
    dem, ref = UHF.Demodulator(conf, P, RADIO), R.RefGpuDemodulator(conf, P, RADIO)
    sig, tx = S.bench_stream(mod, snr, seed=31)
    N, ovl = dem.Nfft, dem.sigOverlap
    step = N - ovl
    rd, rr = dem.get_signalBufferHostPointer(), ref.get_signalBufferHostPointer()
    rd[:] = 0
    rr[:] = 0
    n_sym = n_diff = n_bits = n_bitdiff = 0
    for c in range(len(sig) // step):
        rd[ovl:] = sig[c * step:(c + 1) * step]
        rr[ovl:] = sig[c * step:(c + 1) * step]
        fa, fb = dem.uploadAndFindCarrier(rd), ref.uploadAndFindCarrier(rr)
        ba, bb = dem.demodulate(), ref.demodulate()
        ld, lr = dem.last, ref.last
        assert rel_err(ld["E"], lr["E"]) < 1e-4, f"chunk {c}"
        assert ld["shift"] == lr["shift"], f"chunk {c}"
        assert ld["timing"][0] == lr["timing"][0], f"chunk {c}"
        assert fa[0] == pytest.approx(fb[0], abs=1e-2)
        np.testing.assert_allclose(fa[3], fb[3], rtol=2e-4, atol=2e-4, equal_nan=True)      # SNR
        assert ba[3] == bb[3]
        if c > 0:       # chunk 0 carries the zero-filled overlap (rounding-noise ties, see test_gpu_parity)
            assert len(ld["sym"]) == len(lr["sym"])
            n_sym += len(lr["sym"])
            n_diff += int(np.sum(ld["sym"] != lr["sym"]))
            if len(ba[0]) == len(bb[0]):
                n_bits += len(bb[0])
                n_bitdiff += int(np.sum(ba[0] != bb[0]))
            else:
                n_bitdiff += abs(len(ba[0]) - len(bb[0]))
        rd[:ovl] = rd[-ovl:]
        rr[:ovl] = rr[-ovl:]
    ref.close()
    assert n_diff <= 1e-3 * n_sym, f"{n_diff} of {n_sym} symbol decisions differ"
    assert n_bitdiff <= 1e-3 * n_bits, f"{n_bitdiff} of {n_bits} bits differ"


def test_cc11xx_chunk_against_reference_kernels():
    _need(8, 7, True)
    from oracle.ref_gpu.make_golden_gpu import c1_chunk
    from pycusdr_b200.demodulator import UHF
    conf = conf_variant("CC11xx.json")
    P = protocol_for(conf)
    dem, ref = UHF.Demodulator(conf, P, RADIO), R.RefGpuDemodulator(conf, P, RADIO)
    x = c1_chunk(conf)
    for d in (dem, ref):
        d.get_signalBufferHostPointer()[:] = x
    fa = dem.uploadAndFindCarrier(dem.get_signalBufferHostPointer())
    fb = ref.uploadAndFindCarrier(ref.get_signalBufferHostPointer())
    dem.demodulate()
    ref.demodulate()
    assert rel_err(dem.last["E"], ref.last["E"]) < 1e-4
    assert dem.last["shift"] == ref.last["shift"] and dem.last["timing"][0] == ref.last["timing"][0]
    assert fa[0] == pytest.approx(fb[0], abs=1e-2)
    live = ref.last["mag"] > 1e-6 * ref.last["mag"].max()
    assert np.mean(dem.last["sym"][live] != ref.last["sym"][live]) < 2e-3
    ref.close()
