#!/usr/bin/env python
"""Generate tests/golden/ref_host_*.npz from the reference's OWN Python code.

Run in the authoring container only (``/root/reference`` is not on the GPU box):

    LD_LIBRARY_PATH=/usr/local/cuda/lib64 python oracle/make_golden.py

The reference's host-side code is imported verbatim from ``/root/reference/pyCuSDR`` with three
shims: ``np.float``/``np.int`` aliases (removed in NumPy 1.24), and stub modules
for ``pycuda`` and ``crcmod`` (imported at module scope but not needed by the functions called
here).  No reference source is copied; only inputs and the outputs it produced are stored.

What gets pinned:
  ref_host_filters.npz   protocol.get_filter / get_symbolLUT2 of every shipped protocol
  ref_host_signals.npz   create_signals.py modulators, awgn, get_padded_packet
  ref_host_demod.npz     Demodulator.extractBits / extractBitsNRZs / checkSymbolOverlap /
                         __thresholdInput / computeSNR on recorded inputs
"""
import os
import sys
import types

import numpy as np
import scipy.signal  # noqa: F401  (must be imported before the np.float shim)
import scipy.constants  # noqa: F401

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    np.float = float
    np.int = int
    for name in ("pycuda", "pycuda.driver", "pycuda.compiler", "crcmod"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pycuda.compiler"].SourceModule = object
    sys.modules["pycuda"].driver = sys.modules["pycuda.driver"]
    sys.path.insert(0, os.path.join(REF, "pyCuSDR"))
    sys.path.insert(0, os.path.join(REF, "examples", "benchmark"))
    from protocol.loadProtocol import loadProtocol
    from demodulator.demodulator_base import Demodulator
    import create_signals
    return loadProtocol, Demodulator, create_signals


def main():
    loadProtocol, Demodulator, cs = import_reference()
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.RandomState(20261018)

    # ---- protocol filter banks and LUTs -------------------------------------------------
    conf = {"Main": {"PacketLen": 10000, "RandSeed": 123},
            "Radios": {"Protocol": {"rx_preamble": ["0xaa"], "rx_sync_seq": ["0xd6"], "tx_preamble": ["0xaa"],
                                    "tx_num_preambles": 10, "tx_sync_seq": ["0xd6"]}}}
    filt = {}
    for name, sps, maskSize, nfft in (("bench_GMSK", 16, 3, 1024), ("bench_FSK", 16, 3, 1024),
                                      ("bench_GFSK", 16, 3, 1024), ("bench_BPSK", 16, 5, 1024),
                                      ("CC11xx", 128, 3, 1024), ("bench_GMSK", 8, 4, 512)):
        try:
            P = loadProtocol(name)(conf=conf)
        except Exception as e:       # CC11xx needs lib.shift_registers.PN9 etc.; report, don't hide
            print("protocol", name, "failed to construct:", repr(e))
            raise
        M, masks = P.get_filter(nfft, sps, maskSize)
        key = f"{name}_{sps}_{maskSize}_{nfft}"
        filt[key + "_masks"] = masks
        filt[key + "_M"] = M
        bitLUT, symLUT = P.get_symbolLUT2(maskSize)
        filt[key + "_bitLUT"] = np.array([]) if bitLUT is None else np.asarray(bitLUT)
        filt[key + "_bitLUT_none"] = bitLUT is None
        filt[key + "_symLUT"] = np.asarray(symLUT)
        filt[key + "_sum_all"] = bool(getattr(P, "SUM_ALL_MASKS_PYTHON", False))
    np.savez_compressed(os.path.join(OUT, "ref_host_filters.npz"), **filt)

    # ---- signal generators -----------------------------------------------------------------
    sig = {}
    bits = cs.createBitSequence(96, seed=7)
    sig["bits96_seed7"] = bits
    sig["packet_bits_head"] = cs.packetData()[:256]
    sig["packet_bits_sum"] = int(cs.packetData().sum())
    for mod, fn in (("BPSK", cs.modulateBPSK), ("FSK", cs.modulateFSK), ("GFSK", cs.modulateGFSK2),
                    ("GMSK", cs.modulateGMSK)):
        sig["mod_" + mod] = fn(bits, 16)
    np.random.seed(99)
    sig["awgn_c"] = cs.awgn(sig["mod_GMSK"], 7.5)
    np.random.seed(98)
    sig["awgn_r"] = cs.awgn(np.real(sig["mod_BPSK"]).astype(np.float64), 3.0, measured=False)
    full, raw = cs.get_padded_packet("GMSK", 16, 153600)
    sig["padded_GMSK_len"] = len(full)
    sig["padded_GMSK_head"] = full[9990:10600]
    full, raw = cs.get_padded_packet("FSK", 16, 153600, offset_freq=12345.0, raw_bits=bits)
    sig["padded_FSK_custom"] = full[9000:11600:5]
    sig["nrzs"] = cs.encodeNRZS(bits)
    np.savez_compressed(os.path.join(OUT, "ref_host_signals.npz"), **sig)

    # ---- Demodulator host-side methods ----------------------------------------------------------
    dm = {}
    d = object.__new__(Demodulator)
    d.GPU_buffers, d.GPU_fftPlans = [], []           # __del__ expects them
    # extractBits (bitLUT path), symbols include -1 (all-zero window)
    d.bitLUT = np.array([0., 0., 1., 1., 0., 0., 1., 1.])
    d.symbolLUT = []
    sym = rng.randint(-1, 8, 400).astype(np.int32)
    out, err = d.extractBits(np.arange(400), sym)
    dm["eb_sym"], dm["eb_bits"], dm["eb_nerr"] = sym, out, len(err)
    # extractBitsNRZs with the BPSK maskLen-4 LUT
    Pb = loadProtocol("bench_BPSK")(conf=conf)
    _, lut = Pb.get_symbolLUT2(4)
    d.bitLUT, d.symbolLUT = None, lut
    sym = rng.randint(0, 8, 300).astype(np.int32)
    out, err = d.extractBits(np.arange(300), sym)
    dm["nrzs_lut"], dm["nrzs_sym"], dm["nrzs_bits"], dm["nrzs_err"] = lut, sym, out.astype(np.int8), np.array(err)

    # checkSymbolOverlap: a sequence of chunks with deliberate early / late slips
    N, ovl, sps = 4096, 256, 16
    d.Nfft, d.sigOverlapWin = N, ovl // 2
    d.overlapOffset, d.symbol_check_error_threshold, d.symbol_check_match_threshold = 6, 1000, 3
    d.poswinP = []
    stream = rng.randint(0, 2, 4000).astype(np.float64)
    cases = []
    pos = 0
    for c, slip in enumerate((0, 0, 1, 0, -1, 0, 0, 2, 0)):
        S = N // sps
        first_centre = 5 + (c % 3)
        centres = (first_centre + np.arange(S) * sps).astype(np.int32)
        # symbols covering this chunk: chunk advances by (N-ovl)/sps symbols; 'slip' shifts the window
        start = pos + slip
        bitsC = stream[start:start + S].copy()
        if c == 5:
            bitsC[3] = 1 - bitsC[3]                 # a bit error inside the compared region
        trust = rng.randint(-100, 100, S).astype(np.int8)
        cW, bW, tW, _ = d.checkSymbolOverlap(0, centres, bitsC.astype(np.int32), bitsC, trust)
        cases.append((centres, bitsC, trust, cW, bW, tW))
        pos += (N - ovl) // sps
    for i, (ce, bi, tr, cW, bW, tW) in enumerate(cases):
        dm[f"cso{i}_centres"], dm[f"cso{i}_bits"], dm[f"cso{i}_trust"] = ce, bi, tr
        dm[f"cso{i}_cW"], dm[f"cso{i}_bW"], dm[f"cso{i}_tW"] = cW, bW, tW
    dm["cso_n"] = len(cases)
    dm["cso_params"] = np.array([N, ovl // 2, 6, 1000, 3])

    # __thresholdInput
    d.Nfft = 2048
    d.peakThresholdScale = 4.5
    x = (rng.randn(2048) + 1j * rng.randn(2048)).astype(np.complex64)
    x[100:104] *= 40
    x[130] *= 25
    x[900] *= 300
    x[1500:1503] *= 12
    dm["thr_in"] = x.copy()
    d._Demodulator__thresholdInput(x)
    dm["thr_out"], dm["thr_pure"], dm["thr_filled"] = x, np.asarray(d.clippedPeakIPure), np.asarray(d.clippedPeakI)

    # computeSNR (cuda.Context.synchronize is stubbed)
    drv = sys.modules["pycuda.driver"]
    drv.Context = types.SimpleNamespace(synchronize=lambda: None)
    import demodulator.demodulator_base as dbmod
    dbmod.cuda = drv
    d.Nfft = 4096
    X = (rng.randn(4096) + 1j * rng.randn(4096)).astype(np.complex64)
    X[1000:1040] *= 9
    d.GPU_bufSignalFreq_cpu_handle = X
    d.doppCyperSymNorm = np.array([990, 1010, 1030, 1050, 4090, 6, 2040, 2060], dtype=np.int32)
    snr = [d.computeSNR(lo, hi, 5) for lo, hi in ((0, 1), (1, 2), (2, 3), (4, 5), (6, 7))]
    dm["snr_X"], dm["snr_shifts"], dm["snr_vals"] = X, d.doppCyperSymNorm, np.array(snr, dtype=np.float64)
    dm["snr_pairs"] = np.array([(0, 1), (1, 2), (2, 3), (4, 5), (6, 7)])
    np.savez_compressed(os.path.join(OUT, "ref_host_demod.npz"), **dm)
    print("golden fixtures written to", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
