#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- run on a B200 (``gpurun -- python -m oracle.ref_gpu.make_golden_gpu``).

Feeds seeded synthetic chunks through the reference's own device code (``oracle/ref_gpu/driver.py``: the unmodified
``cuda_kernels.cu`` + cuFFT, launched exactly as ``demodulator_base.py`` does) and stores what it produced in
``gpurun_out/refgpu_*.npz``; the files are then committed under ``tests/golden/``.  They are the fixtures that pin
the oracle's restatement of the eight per-chunk kernels: ``tests/test_oracle_refgpu.py`` replays
the same seeded inputs through ``oracle/oracle.py`` on the CPU and compares.

  refgpu_streams.npz   per chunk of several seeded streams / protocols / modes: E[D,M], findDopplerEst result,
                       selected shift, findCodeRateAndPhase result, symbols, centres, magnitudes, output bits
  refgpu_small.npz     one small chunk with every intermediate buffer (spectrum, full correlation surface,
                       demod surface, summed power, its R2C spectrum) so that each restated kernel can be checked
                       on the reference's own inputs, bit for bit where the arithmetic is exact
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import signals as S                                    # noqa: E402
from oracle.ref_gpu.driver import RefGpuDemodulator                # noqa: E402
from tests.helpers import RADIO, conf_variant, protocol_for        # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")

# name: (config, blockSize, modulation, bench SNR, seed, chunks kept, overrides)
STREAM_CASES = {
    "gmsk20": ("benchmark/bench_GMSK.json", 15, "GMSK", 20, 11, 8, {}),
    "gmsk9": ("benchmark/bench_GMSK.json", 15, "GMSK", 9, 12, 8, {}),
    "fsk14": ("benchmark/bench_FSK.json", 15, "FSK", 14, 13, 8, {}),
    "gfsk14": ("benchmark/bench_GFSK.json", 15, "GFSK", 14, 14, 8, {}),
    "bpsk14": ("benchmark/bench_BPSK.json", 15, "BPSK", 14, 15, 8, {}),
    "gmsk_nosum": ("benchmark/bench_GMSK.json", 14, "GMSK", 15, 16, 6, {"sum_all": False}),
    "gmsk_noise_row": ("benchmark/bench_GMSK.json", 14, "GMSK", 15, 17, 6, {"noise_measure_offset_Hz": 60000.0}),
    "gmsk_stx": ("benchmark/bench_GMSK.json", 14, "GMSK", 15, 18, 6, {"backend": "STX", "peakThresholdScale": 4.5}),
}


def stream_for(case):
    cfg, blockSize, mod, snr, seed, keep, over = STREAM_CASES[case]
    radio_over = {k: v for k, v in over.items() if k == "noise_measure_offset_Hz"}
    conf = conf_variant(cfg, blockSize=blockSize, **radio_over)
    if "peakThresholdScale" in over:
        conf["GPU"]["UHF"]["peakThresholdScale"] = over["peakThresholdScale"]
    P = protocol_for(conf)
    if "sum_all" in over:
        P.SUM_ALL_MASKS_PYTHON = over["sum_all"]
    sig, bits = S.bench_stream(mod, snr, seed=seed)
    if over.get("backend") == "STX":
        sig = sig.copy()
        sig[30000:30003] *= 60
    return conf, P, sig, keep, over.get("backend", "UHF")


def c1_chunk(conf):
    """FSK-2 packet of 400 bits at sample 4096, +7 kHz off the nominal offset, Es/N0 15 dB."""
    cr = conf["Radios"]["Rx"][RADIO]
    sps, fs = cr["samplesPerSym"], cr["baud"] * cr["samplesPerSym"]
    N = 2 ** conf["GPU"]["UHF"]["blockSize"]
    bits = S.createBitSequence(400, seed=123)
    sig = S.modulateFSK(bits, sps)
    x = np.zeros(N, dtype=np.complex128)
    x[4096:4096 + len(sig)] = sig[:N - 4096]
    x *= np.exp(2j * np.pi * (cr["frequencyOffset_Hz"] + 7000.0) / fs * np.arange(N))
    rng = np.random.RandomState(1)
    amp = np.sqrt(10 ** (-(15 - 10 * np.log10(sps)) / 10) / 2)
    return (x + amp * (rng.randn(N) + 1j * rng.randn(N))).astype(np.complex64)


def main():
    os.makedirs(OUT, exist_ok=True)
    out = {}
    for case in STREAM_CASES:
        conf, P, sig, keep, backend = stream_for(case)
        dem = RefGpuDemodulator(conf, P, RADIO, backend=backend)
        N, ovl = dem.Nfft, dem.sigOverlap
        step = N - ovl
        raw = dem.get_signalBufferHostPointer()
        raw[:] = 0
        first = 4 if backend == "UHF" else 1        # skip the noise-only lead-in of the bench stream
        for c in range(len(sig) // step):
            raw[ovl:] = sig[c * step:(c + 1) * step]
            f = dem.uploadAndFindCarrier(raw)
            b = dem.demodulate()
            if first <= c < first + keep:
                k = f"{case}/{c}/"
                L = dem.last
                out[k + "xsum"] = np.array([np.sum(raw.astype(np.complex128))])
                if backend == "UHF":
                    out[k + "E"] = L["E"]
                    out[k + "res"] = L["res"]
                    out[k + "ret"] = np.array([f[0], f[1], f[3]], dtype=np.float64)
                out[k + "shift"] = np.array([L["shift"]], dtype=np.int64)
                out[k + "timing"] = L["timing"]
                out[k + "sym"] = L["sym"].astype(np.int16)
                out[k + "centres"] = L["centres"]
                out[k + "mag"] = L["mag"]
                out[k + "bits"] = np.packbits(b[0])
                out[k + "nbits"] = np.array([len(b[0])])
                out[k + "trust"] = b[2]
                out[k + "clipped"] = np.asarray(dem.clippedPeakIPure, dtype=np.int64)
            raw[:ovl] = raw[-ovl:]
        print(case, "done,", dem.launches, "launches")
        dem.close()
    # C1 plumbing chunk (CC11xx, N = 2^16)
    conf = conf_variant("CC11xx.json")
    P = protocol_for(conf)
    dem = RefGpuDemodulator(conf, P, RADIO)
    x = c1_chunk(conf)
    dem.get_signalBufferHostPointer()[:] = x
    f = dem.uploadAndFindCarrier(dem.get_signalBufferHostPointer())
    b = dem.demodulate()
    L = dem.last
    out.update({"c1/0/xsum": np.array([np.sum(x.astype(np.complex128))]), "c1/0/E": L["E"], "c1/0/res": L["res"],
                "c1/0/ret": np.array([f[0], f[1], f[3]]), "c1/0/shift": np.array([L["shift"]], dtype=np.int64),
                "c1/0/timing": L["timing"], "c1/0/sym": L["sym"].astype(np.int16), "c1/0/centres": L["centres"],
                "c1/0/mag": L["mag"], "c1/0/bits": np.packbits(b[0]), "c1/0/nbits": np.array([len(b[0])]),
                "c1/0/trust": b[2], "c1/0/clipped": np.zeros(0, dtype=np.int64)})
    dem.close()
    np.savez_compressed(os.path.join(OUT, "refgpu_streams.npz"), **out)

    # ---- one small chunk with all intermediates ----
    conf = conf_variant("benchmark/bench_GMSK.json", blockSize=12, doppCarrierSteps=6)
    P = protocol_for(conf)
    P.SUM_ALL_MASKS_PYTHON = False
    dem = RefGpuDemodulator(conf, P, RADIO)
    N, D, M = dem.Nfft, dem.doppIdxArrayLen, dem.num_masks
    sig, _ = S.get_padded_packet("GMSK", 16, 153600, pad=300, raw_bits=S.createBitSequence(230, seed=5))
    rng = np.random.RandomState(77)
    x = (sig[:N] + 0.2 * (rng.randn(N) + 1j * rng.randn(N))).astype(np.complex64)
    dem.get_signalBufferHostPointer()[:] = x
    dem.uploadToGPU()
    best = dem.search_device()
    surf = np.empty((D, M, N), dtype=np.complex64)
    dem._dtoh(surf, dem.bufXcorr)
    E = dem.energies()
    dem.sync()
    X = dem.sigFreq_host.copy()
    lo, hi, hz, shift = dem_interp(dem, best)
    r, spSym, codeOffset, sym, centres, mag = dem.demod_device(int(shift))
    y = dem.demod_surface()
    p = np.empty(N, dtype=np.float32)
    dem._dtoh(p, dem.bufCodeAndPhase)
    Pf = np.empty(N // 2 + 1, dtype=np.complex64)
    dem._dtoh(Pf, dem.bufCodeAndPhaseOut)
    np.savez_compressed(os.path.join(OUT, "refgpu_small.npz"), x=x, X=X, surface=surf, E=E, res=best,
                        shift=np.array([shift]), y=y, p=p, Pf=Pf, timing=r, sym=sym, centres=centres, mag=mag,
                        shifts=dem.doppCyperSymNorm)
    dem.close()
    print("wrote refgpu_streams.npz, refgpu_small.npz")


def dem_interp(dem, best):
    from oracle import oracle as O
    return O.interpolate_doppler(best[0], dem.doppCyperSymNorm, dem.doppHzLUT)


if __name__ == "__main__":
    main()
