"""GPU parity tests proper: the CUDA path (through the C ABI) against the NumPy oracle on the same
seeded inputs.  Tolerances (BASELINE.json north_star / SURVEY.md A.2):
  * spectrum, correlation surface, energies E[D,M], peaks: <= 1e-4 relative (fp32);
  * selected Doppler bin / spectrum shift, timing bin: identical (ties: lowest index);
  * timing phase: <= 1e-4 rad;
  * symbol decisions, centres, magnitudes: bit-exact given identical surface and timing inputs;
  * demodulated bits: bit-exact per chunk.
"""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import signals as S
from tests.helpers import RADIO, conf_variant, load_conf, protocol_for, rel_err

pytestmark = pytest.mark.gpu


def _demods(conf, **kw):
    from pycusdr_b200.demodulator import UHF
    P = protocol_for(conf)
    return UHF.Demodulator(conf, P, RADIO, **kw), O.OracleDemodulator(conf, P, RADIO)


def _noise_chunk(N, seed, with_packet=None, sps=16):
    rng = np.random.RandomState(seed)
    x = (rng.randn(N) + 1j * rng.randn(N)).astype(np.complex64) * np.float32(0.5)
    if with_packet is not None:
        sig, _ = S.get_padded_packet(with_packet, sps, 9600 * sps, pad=100)
        n = min(len(sig), N - 200)
        x[100:100 + n] += sig[:n].astype(np.complex64)
    return x


@pytest.mark.parametrize("blockSize", [12, 15, 16, 18])
def test_chunk_spectrum_matches_fft(blockSize):
    conf = conf_variant("benchmark/bench_GMSK.json", blockSize=blockSize, doppCarrierSteps=4)
    dem, _ = _demods(conf)
    x = _noise_chunk(2 ** blockSize, 3)
    dem.get_signalBufferHostPointer()[:] = x
    dem.uploadToGPU(dem.get_signalBufferHostPointer())
    X = dem._engine.spectrum()
    ref = np.fft.fft(x.astype(np.complex128))
    assert rel_err(X, ref) < 2e-6


@pytest.mark.parametrize("cfg,blockSize,log2_block", [
    ("benchmark/bench_GMSK.json", 15, 0), ("benchmark/bench_GMSK.json", 15, 8), ("benchmark/bench_GMSK.json", 15, 9),
    ("benchmark/bench_FSK.json", 12, 8), ("benchmark/bench_GMSK.json", 15, 11),
    ("benchmark/bench_GMSK.json", 15, 12), ("benchmark/bench_FSK.json", 14, 10), ("CC11xx.json", 16, 0),
    ("CC11xx.json", 16, 13), ("benchmark/bench_BPSK.json", 13, 0)])
def test_search_energy_and_peaks_match_oracle(cfg, blockSize, log2_block):
    conf = conf_variant(cfg, blockSize=blockSize)
    dem, orc = _demods(conf, fused=False, log2_block=log2_block)
    sps = conf["Radios"]["Rx"][RADIO]["samplesPerSym"]
    N = 2 ** blockSize
    x = _noise_chunk(N, 5, with_packet="GMSK" if sps == 16 else None, sps=sps)
    dem.get_signalBufferHostPointer()[:] = x
    dem.uploadToGPU(dem.get_signalBufferHostPointer())
    res, E = dem._engine.search()
    X = O.forward_fft(x)
    Eo, pv, po = O.search_energy(X, orc.masks, orc.doppCyperSymNorm, orc.SUM_ALL_MASKS_PYTHON, want_peaks=True)
    assert rel_err(E, Eo) < 1e-4
    ro = O.find_doppler_est(Eo, orc.num_dopplers, orc.doppIdxArrayOffset, orc.SUM_ALL_MASKS_PYTHON)
    assert abs(res.best_idx - ro[0]) < 1e-3
    assert abs(res.metric_db - ro[1]) < 1e-3
    # feeding the oracle's estimator with the CUDA energies must reproduce the CUDA estimate bit for bit
    rc = O.find_doppler_est(E, orc.num_dopplers, orc.doppIdxArrayOffset, orc.SUM_ALL_MASKS_PYTHON)
    assert np.float32(res.best_idx) == rc[0]
    lo, hi, hz, shift = O.interpolate_doppler(np.float32(res.best_idx), orc.doppCyperSymNorm, orc.doppHzLUT)
    assert (res.low_idx, res.high_idx, res.shift) == (lo, hi, int(shift))
    v, o = dem._engine.peaks()
    assert rel_err(v, pv) < 1e-4
    # offsets must agree wherever the oracle's maximum is unambiguous at the 1e-4 level
    y_amb = 0
    for d in range(v.shape[0]):
        for m in range(v.shape[1]):
            if o[d, m] != po[d, m]:
                y_amb += 1
    assert y_amb <= 0.02 * v.size
    gv = np.unravel_index(np.argmax(pv), pv.shape)
    assert (res.peak_bin, res.peak_mask) == gv and res.peak_offset == po[gv]


@pytest.mark.parametrize("cfg,blockSize", [("benchmark/bench_GMSK.json", 15), ("CC11xx.json", 16)])
def test_demod_surface_timing_and_centres(cfg, blockSize):
    conf = conf_variant(cfg, blockSize=blockSize)
    dem, orc = _demods(conf, fused=False)
    sps = conf["Radios"]["Rx"][RADIO]["samplesPerSym"]
    N = 2 ** blockSize
    x = _noise_chunk(N, 7, with_packet="GMSK" if sps == 16 else None, sps=sps)
    dem.get_signalBufferHostPointer()[:] = x
    dem.uploadToGPU(dem.get_signalBufferHostPointer())
    shift = int(orc.doppCyperSymNorm[len(orc.doppCyperSymNorm) // 2]) + 3
    X = O.forward_fft(x)
    yo = O.surface_rows(X, orc.masks, shift)
    y = dem._engine.demod_surface(shift)
    assert np.max(np.abs(y - yo)) < 1e-4 * np.max(np.abs(yo))
    res, sym, centre, mag = dem._engine.demod(shift)
    ymag, p = dem._engine.demod_magnitudes()
    assert rel_err(ymag, O.abs2(yo)) < 1e-4
    po = O.sum_masks_abs2(yo)
    assert rel_err(p, po) < 1e-4
    # timing: same bin, phase within 1e-4 rad
    Pf = np.fft.rfft(po.astype(np.float64))
    ro = O.find_code_rate_and_phase(Pf.astype(np.complex64), orc.iHigh, orc.iLow - orc.iHigh)
    assert res.timing[0] == ro[0]
    dphi = np.angle(np.exp(1j * (float(res.timing[1]) - float(ro[1]))))
    assert abs(dphi) < 1e-4
    spSym, off = O.code_rate_host(np.array(res.timing[:], dtype=np.float32), N)
    assert res.sp_sym == spSym and res.code_offset == off
    # symbol decisions: bit-exact when the oracle is given the very same magnitudes and timing
    so, co, mo = O.find_centres(ymag, spSym, off, N, orc.windowWidth, orc.spsymMin)
    assert res.n_sym == len(so)
    np.testing.assert_array_equal(sym, so)
    np.testing.assert_array_equal(centre, co)
    np.testing.assert_array_equal(mag, mo)


def _run_both(dem, orc, sig):
    """The reference's chunk loop (demodulator_process.py:284-338) on both implementations, keeping the raw
    per-chunk device outputs (``.last``) for symbol-level comparison."""
    N, ovl = dem.Nfft, dem.sigOverlap
    step = N - ovl
    rd, ro = dem.get_signalBufferHostPointer(), orc.get_signalBufferHostPointer()
    rd[:] = 0
    ro[:] = 0
    out = []
    for c in range(len(sig) // step):
        rd[ovl:] = sig[c * step:(c + 1) * step]
        ro[ovl:] = sig[c * step:(c + 1) * step]
        fa, fb = dem.uploadAndFindCarrier(rd), orc.uploadAndFindCarrier(ro)
        ba, bb = dem.demodulate(), orc.demodulate()
        out.append((fa, fb, ba, bb, dict(dem.last), dict(orc.last)))
        rd[:ovl] = rd[-ovl:]
        ro[:ovl] = ro[-ovl:]
    return out


@pytest.mark.parametrize("mod,cfg,snr", [("GMSK", "benchmark/bench_GMSK.json", 20), ("GMSK", "benchmark/bench_GMSK.json", 10),
                                         ("FSK", "benchmark/bench_FSK.json", 14), ("GFSK", "benchmark/bench_GFSK.json", 14)])
def test_stream_bits_match_oracle_per_chunk(mod, cfg, snr):
    """Identical Doppler shift and timing bin on every chunk, identical symbol decisions wherever the
    correlation is not numerically zero, identical output bits.  Symbols whose whole window is rounding noise
    of an all-zero input stretch (the zero-initialised overlap of the very first chunk) are exempt: there the
    decision is an arg-max over rounding errors in any implementation, the reference's included."""
    conf = load_conf(cfg)
    dem, orc = _demods(conf)
    sig, bits = S.bench_stream(mod, snr, seed=11)
    chunks = _run_both(dem, orc, sig)
    assert len(chunks) > 5
    allbits = []
    for c, (fa, fb, ba, bb, ld, lo) in enumerate(chunks):
        assert ld["shift"] == lo["shift"], f"chunk {c}: spectrum shift"
        assert ld["timing"][0] == lo["timing"][0], f"chunk {c}: timing bin"
        assert abs(np.angle(np.exp(1j * (float(ld["timing"][1]) - float(lo["timing"][1]))))) < 1e-4
        assert fa[0] == pytest.approx(fb[0], abs=1e-2)              # frequency offset [Hz]
        assert ba[3] == bb[3]                                       # spSym
        live = lo["mag"] > 1e-9 * lo["mag"].max()
        if c == 0:
            # the filters' leading edge over the zero-filled overlap sees 1..L-1 samples only; constant-envelope
            # (FSK) filters tie exactly there, so those decisions are rounding noise as well
            edge = 2 * dem.spsym * conf["GPU"]["UHF"]["xcorrMaskSize"]
            live &= lo["centres"] >= dem.sigOverlap + edge
            # ... and the circular correlation wraps the chunk's last L-1 outputs onto that same zero stretch
            live &= lo["centres"] < dem.Nfft - edge
        assert len(ld["sym"]) == len(lo["sym"])
        np.testing.assert_array_equal(ld["sym"][live], lo["sym"][live], err_msg=f"chunk {c}: symbols")
        np.testing.assert_array_equal(ld["centres"][live], lo["centres"][live], err_msg=f"chunk {c}: centres")
        if live.all():
            np.testing.assert_array_equal(ba[0], bb[0], err_msg=f"chunk {c}: bits")
        else:
            assert c == 0, "only the first chunk has a zero-filled overlap"
        allbits.append(ba[0])
    if snr >= 20:
        allbits = np.concatenate(allbits)
        L = len(bits)
        errs = min(int(np.sum(allbits[o:o + L] != bits)) for o in range(len(allbits) - L))
        assert errs == 0


class _HostOracle:
    """The oracle's host-side post-processing (bit extraction, stitching, clip tagging, output casts) driven by
    the CUDA path's own device outputs: the product's host code must reproduce it exactly, trust bytes included."""

    def __init__(self, orc):
        self.orc, self.state = orc, O.OverlapState()

    def __call__(self, last, clipped):
        o = self.orc
        trust = O.trust_from_magnitudes(last["mag"], len(last["sym"]))
        bits, err = O.extract_bits(last["sym"], o.bitLUT, o.symbolLUT)
        cW, bW, tW = O.check_symbol_overlap(self.state, len(err), last["centres"], bits, trust, o.Nfft, o.sigOverlapWin,
                                            o.overlapOffset, o.symbol_check_error_threshold, o.symbol_check_match_threshold)
        tW = O.tag_clipped_peaks(tW, cW, clipped, last["spSym"], o.Nfft)
        return bW.astype(np.uint8), cW.astype(np.uint8), tW.astype(np.uint8)


@pytest.mark.parametrize("backend", ["UHF", "STX"])
def test_host_postprocessing_exact_on_device_outputs(backend):
    from pycusdr_b200 import demodulator
    conf = conf_variant("benchmark/bench_GMSK.json", blockSize=15)
    conf["GPU"]["UHF"]["peakThresholdScale"] = 4.5
    P = protocol_for(conf)
    dem = getattr(demodulator, backend).Demodulator(conf, P, RADIO)
    host = _HostOracle(O.OracleDemodulator(conf, P, RADIO, backend=backend))
    sig, _ = S.bench_stream("GMSK", 9, seed=21)
    sig = sig.copy()
    sig[50000:50004] *= 80
    N, ovl = dem.Nfft, dem.sigOverlap
    raw = dem.get_signalBufferHostPointer()
    raw[:] = 0
    tagged = 0
    for c in range(len(sig) // (N - ovl)):
        raw[ovl:] = sig[c * (N - ovl):(c + 1) * (N - ovl)]
        dem.uploadAndFindCarrier(raw)
        bits, centres, trust, spSym = dem.demodulate()
        eb, ec, et = host(dem.last, dem.clippedPeakIPure)
        np.testing.assert_array_equal(bits, eb)
        np.testing.assert_array_equal(centres, ec)
        np.testing.assert_array_equal(trust, et)
        tagged += len(dem.clippedPeakIPure)
        raw[:ovl] = raw[-ovl:]
    assert (tagged > 0) == (backend == "STX")


@pytest.mark.parametrize("kw", [dict(search_form=2), dict(search_form=1), dict(items_per_cta=5), dict(items_per_cta=1000),
                                dict(groups_per_cta=4, items_per_cta=7), dict(groups_per_cta=16, items_per_cta=33)])
def test_search_forms_agree(kw):
    """The shifted-filter form of the 256-point search (default), the rotate-the-chunk forms and every CTA tiling of the
    former give the same energies and peaks (<= 2e-6: only fp32 rounding differs) and the same estimate; tilings of the
    same form are bit-identical because each (bin, block) partial is computed by the same instruction sequence."""
    conf = conf_variant("benchmark/bench_GMSK.json", blockSize=14, doppCarrierSteps=24,
                        noise_measure_offset_Hz=30000)          # noise row prepended: 25 rows
    demA, orc = _demods(conf, fused=False)
    demB, _ = _demods(conf, fused=False, **kw)
    x = _noise_chunk(2 ** 14, 11, with_packet="GMSK")
    out = []
    for dem in (demA, demB):
        dem.get_signalBufferHostPointer()[:] = x
        dem.uploadToGPU(dem.get_signalBufferHostPointer())
        res, E = dem._engine.search()
        v, o = dem._engine.peaks()
        out.append((res, E.copy(), v.copy(), o.copy()))
    (ra, Ea, va, oa), (rb, Eb, vb, ob) = out
    if "search_form" in kw:
        assert rel_err(Eb, Ea) < 2e-6 and rel_err(vb, va) < 2e-6
        assert np.mean(oa != ob) <= 0.02
    else:
        assert np.array_equal(Ea, Eb) and np.array_equal(va, vb) and np.array_equal(oa, ob)
    assert (ra.shift, ra.low_idx, ra.high_idx) == (rb.shift, rb.low_idx, rb.high_idx)
    Eo, pv, po = O.search_energy(O.forward_fft(x), orc.masks, orc.doppCyperSymNorm, orc.SUM_ALL_MASKS_PYTHON, want_peaks=True)
    assert rel_err(Eb, Eo) < 1e-4 and rel_err(vb, pv) < 1e-4


@pytest.mark.parametrize("log2_block", [9, 11])
def test_generic_kernel_forms_agree(log2_block):
    """Shifted-filter and rotate-the-chunk forms of the generic (shared-memory) search kernel: same numbers to fp32 rounding."""
    conf = conf_variant("benchmark/bench_GMSK.json", blockSize=14, doppCarrierSteps=24, noise_measure_offset_Hz=30000)
    demA, orc = _demods(conf, fused=False, log2_block=log2_block)
    demB, _ = _demods(conf, fused=False, log2_block=log2_block, search_form=2)
    x = _noise_chunk(2 ** 14, 13, with_packet="GMSK")
    out = []
    for dem in (demA, demB):
        dem.get_signalBufferHostPointer()[:] = x
        dem.uploadToGPU(dem.get_signalBufferHostPointer())
        res, E = dem._engine.search()
        v, o = dem._engine.peaks()
        out.append((res, E.copy(), v.copy(), o.copy()))
    (ra, Ea, va, oa), (rb, Eb, vb, ob) = out
    assert demA._engine.launch_count == demB._engine.launch_count + 1        # the block-spectra kernel
    assert rel_err(Ea, Eb) < 2e-6 and rel_err(va, vb) < 2e-6 and np.mean(oa != ob) <= 0.02
    assert (ra.shift, ra.low_idx, ra.high_idx) == (rb.shift, rb.low_idx, rb.high_idx)


@pytest.mark.parametrize("cfg,blockSize,log2_block,sjr", [("CC11xx.json", 16, 0, (128, 3, 2)), ("CC11xx.json", 14, 12, (128, 3, 2)),
                                                         ("benchmark/bench_FSK.json", 14, 10, (16, 3, 2))])
def test_factorised_bank_matches_the_unfactorised_search(cfg, blockSize, log2_block, sjr):
    """search_fb_kernel (R = 2 transforms per item + J-term combinations; bank_factor.cu) against search_os_kernel on the
    unfactorised per-bin spectra (search_form = 3): same energies and peaks to fp32 rounding, same estimate, and against
    the oracle <= 1e-4.  A slice of the bins (sharding) is bit-identical to the same rows of the full search."""
    conf = conf_variant(cfg, blockSize=blockSize, doppCarrierSteps=24, noise_measure_offset_Hz=30000)
    demA, orc = _demods(conf, fused=False, log2_block=log2_block)
    demB, _ = _demods(conf, fused=False, log2_block=log2_block, search_form=3)
    assert demA._engine.bank_factor() == (3,) + sjr and demB._engine.bank_factor()[0] == 0      # FSK-2: complete, shared prefixes
    sps = conf["Radios"]["Rx"][RADIO]["samplesPerSym"]
    x = _noise_chunk(2 ** blockSize, 17, with_packet="FSK" if sps == 16 else None, sps=sps)
    if sps != 16:       # CC11xx: an FSK-2 burst at the radio's offset so that the tables are not noise only
        bits = S.createBitSequence(60, seed=5)
        sig = S.modulateFSK(bits, sps)[: 2 ** blockSize - 512].astype(np.complex64)
        fs = conf["Radios"]["Rx"][RADIO]["baud"] * sps
        f0 = conf["Radios"]["Rx"][RADIO]["frequencyOffset_Hz"] + 3000.0
        x[256:256 + len(sig)] += 2 * sig * np.exp(2j * np.pi * f0 / fs * np.arange(len(sig))).astype(np.complex64)
    out = []
    for dem in (demA, demB):
        dem.get_signalBufferHostPointer()[:] = x
        dem.uploadToGPU(dem.get_signalBufferHostPointer())
        res, E = dem._engine.search()
        v, o = dem._engine.peaks()
        out.append((res, E.copy(), v.copy(), o.copy()))
    (ra, Ea, va, oa), (rb, Eb, vb, ob) = out
    assert rel_err(Ea, Eb) < 5e-6 and rel_err(va, vb) < 5e-6 and np.mean(oa != ob) <= 0.02
    assert (ra.shift, ra.low_idx, ra.high_idx) == (rb.shift, rb.low_idx, rb.high_idx)
    Eo, pv, po = O.search_energy(O.forward_fft(x), orc.masks, orc.doppCyperSymNorm, orc.SUM_ALL_MASKS_PYTHON, want_peaks=True)
    assert rel_err(Ea, Eo) < 1e-4 and rel_err(va, pv) < 1e-4
    # bin slices: rows [lo, hi) on their own == the same rows of the full table
    eng = demA._engine
    import torch
    dev = []
    for ptr, typestr in zip(eng.shard_buffers(), ("<f4", "<f4", "<i4")):
        class _W:
            __cuda_array_interface__ = {"shape": (eng.D, eng.M), "typestr": typestr, "data": (ptr, False), "version": 2}
        dev.append(torch.as_tensor(_W(), device="cuda:0"))

    def tables():
        torch.cuda.synchronize()
        got = [t.cpu().numpy().copy() for t in dev]
        for t in dev:
            t.zero_()
        torch.cuda.synchronize()
        return got

    eng.upload()
    eng.enqueue_search_local()
    E, v, o = tables()
    for lo, hi in ((0, 9), (9, 25)):
        eng.set_bin_range(lo, hi)
        eng.upload()
        eng.enqueue_search_local()
        Es, vs, os_ = tables()
        np.testing.assert_array_equal(Es[lo:hi], E[lo:hi])
        np.testing.assert_array_equal(vs[lo:hi], v[lo:hi])
        np.testing.assert_array_equal(os_[lo:hi], o[lo:hi])


@pytest.mark.parametrize("bank", ["incomplete", "three_tones", "two_segments", "two_segments_no_tree", "cc11xx_no_tree"])
def test_factorised_bank_general_form(bank):
    """Banks that factorise but are not a complete binary bank take search_fb_kernel's general epilogue (selectors read at
    run time): 7 of the 8 CC11xx templates; 9 two-segment templates over three tones (R = 3); the 4 two-symbol FSK-2
    templates (complete, J = 2).  Against the unfactorised search on the same spectra, and against the oracle."""
    from pycusdr_b200 import _native
    from pycusdr_b200.protocol.FSK2_base import fsk_phase_templates
    N, sps = 2 ** 14, 128
    if bank == "incomplete":
        pats = [np.array([(k >> 2) & 1, (k >> 1) & 1, k & 1]) for k in range(8) if k != 5]
        tmpl = fsk_phase_templates(pats, sps, 0.5)
        want = (1, 128, 3, 2)
    elif bank.startswith("two_segments"):
        tmpl = fsk_phase_templates([np.array([a, b]) for a in (0, 1) for b in (0, 1)], sps, 0.5)
        want = (3 if bank == "two_segments" else 2, 128, 2, 2)
    elif bank == "cc11xx_no_tree":
        tmpl = fsk_phase_templates([np.array([(k >> 2) & 1, (k >> 1) & 1, k & 1]) for k in range(8)], sps, 0.5)
        want = (2, 128, 3, 2)
    else:
        n = np.arange(sps)
        tones = [np.exp(2j * np.pi * f * n / sps) for f in (-1.0, 0.5, 1.5)]
        tmpl = [np.concatenate((tones[a], np.exp(0.3j * (a + 2 * b)) * tones[b])) for a in range(3) for b in range(3)]
        want = (1, 128, 2, 3)
    M = len(tmpl)
    masks = np.zeros((M, N), dtype=np.complex128)
    for m, tp in enumerate(tmpl):
        masks[m, :len(tp)] = tp
    masks = np.conj(np.fft.fft(masks, axis=1)).astype(np.complex64)        # protocolBase._pad_and_conj_fft
    D = 12
    shifts = ((np.arange(D) - D // 2) * 37) % N
    kw = dict(device=0, nfft=N, num_dopplers=D, element_offset=0, shifts=shifts, masks=masks, window_width=7,
              sum_all_masks=True, code_search_mask_offset=0, samples_per_sym=sps)
    engA, engB = _native.Engine(fb_tree=not bank.endswith("no_tree"), **kw), _native.Engine(search_form=3, **kw)
    assert engA.bank_factor() == want and engB.bank_factor()[0] == 0
    rng = np.random.RandomState(23)
    x = ((rng.randn(N) + 1j * rng.randn(N)) * 0.5).astype(np.complex64)
    x[300:300 + len(tmpl[2])] += (3 * tmpl[2] * np.exp(2j * np.pi * 37 * 2 * np.arange(len(tmpl[2])) / N)).astype(np.complex64)
    out = []
    for eng in (engA, engB):
        eng.host_buffer[:] = x
        eng.upload()
        res, E = eng.search()
        v, o = eng.peaks()
        out.append((E.copy(), v.copy(), o.copy()))
    (Ea, va, oa), (Eb, vb, ob) = out
    assert rel_err(Ea, Eb) < 5e-6 and rel_err(va, vb) < 5e-6 and np.mean(oa != ob) <= 0.02
    Eo, pv, po = O.search_energy(O.forward_fft(x), masks, shifts, True, want_peaks=True)
    assert rel_err(Ea, Eo) < 1e-4 and rel_err(va, pv) < 1e-4
    engA.close()
    engB.close()


@pytest.mark.parametrize("log2_block", [0, 10])
def test_search_bin_range_rows_equal_the_full_table(log2_block):
    """Bin sharding: rows [lo, hi) searched on their own are bit-identical to the same rows of the full search."""
    conf = conf_variant("benchmark/bench_GMSK.json", blockSize=14, doppCarrierSteps=24)
    dem, _ = _demods(conf, fused=False, log2_block=log2_block)
    eng = dem._engine
    x = _noise_chunk(2 ** 14, 12, with_packet="GMSK")
    import torch

    dev = []
    for ptr, typestr in zip(eng.shard_buffers(), ("<f4", "<f4", "<i4")):
        class _W:
            __cuda_array_interface__ = {"shape": (eng.D, eng.M), "typestr": typestr, "data": (ptr, False), "version": 2}
        dev.append(torch.as_tensor(_W(), device="cuda:0"))

    def tables():
        torch.cuda.synchronize()
        out = [t.cpu().numpy().copy() for t in dev]
        for t in dev:
            t.zero_()             # the next run must write its rows itself
        torch.cuda.synchronize()
        return out

    dem.get_signalBufferHostPointer()[:] = x
    eng.upload()
    eng.enqueue_search_local()
    E, v, o = tables()
    assert np.all(E[:, 0] > 0)
    for lo, hi in ((0, 7), (7, 19), (19, 24)):
        eng.set_bin_range(lo, hi)
        eng.upload()
        eng.enqueue_search_local()
        Es, vs, os_ = tables()
        np.testing.assert_array_equal(Es[lo:hi], E[lo:hi])
        np.testing.assert_array_equal(vs[lo:hi], v[lo:hi])
        np.testing.assert_array_equal(os_[lo:hi], o[lo:hi])


@pytest.mark.parametrize("mod,cfg", [("GMSK", "benchmark/bench_GMSK.json"), ("BPSK", "benchmark/bench_BPSK.json")])
def test_fused_and_two_step_schedules_agree(mod, cfg):
    """One native call per chunk (pcs_chunk_to_bits, default), fused device schedule + post-processing in demodulate(),
    the reference's two-step schedule, and the NumPy post-processing mirror: identical outputs."""
    conf = load_conf(cfg)
    demA, _ = _demods(conf)
    assert demA.one_call and demA.fused and demA._stitch is not None
    others = [_demods(conf, one_call=False)[0], _demods(conf, fused=False)[0], _demods(conf, native_post=False)[0]]
    sig, _ = S.bench_stream(mod, 12, seed=5)
    a = O.run_stream(demA, sig)
    assert len(a) > 5
    for dem in others:
        for x, y in zip(a, O.run_stream(dem, sig)):
            np.testing.assert_array_equal(x["data"], y["data"])
            np.testing.assert_array_equal(x["trust"], y["trust"])
            np.testing.assert_equal(x["doppler"], y["doppler"])
            np.testing.assert_equal(x["doppler_std"], y["doppler_std"])
            np.testing.assert_equal(x["SNR"], y["SNR"])       # NaN-aware
            assert x["spSymEst"] == y["spSymEst"]


@pytest.mark.parametrize("cfg,blockSize,reps", [("benchmark/bench_GMSK.json", 15, 3), ("CC11xx.json", 16, 12)])
def test_search_is_bit_reproducible(cfg, blockSize, reps):
    """Fixed-order reductions: the energy table and the peaks are identical run to run (the reference's float atomics are not,
    kern:463,474).  For CC11xx (factorised bank: search_fb_kernel rotates two transforms through three aliased shared-memory
    buffers) this doubles as the race check compute-sanitizer cannot give on this pool: a missing barrier shows up as a
    result that depends on warp timing, and the twelve launches run next to a second handle's searches on another stream."""
    conf = conf_variant(cfg, blockSize=blockSize)
    dem, _ = _demods(conf, fused=False)
    other, _ = _demods(conf, fused=False)
    sps = conf["Radios"]["Rx"][RADIO]["samplesPerSym"]
    x = _noise_chunk(2 ** blockSize, 9, with_packet="GMSK" if sps == 16 else None, sps=sps)
    dem.get_signalBufferHostPointer()[:] = x
    other.get_signalBufferHostPointer()[:] = x[::-1]
    out = []
    for _ in range(reps):
        other._engine.upload()
        other._engine.enqueue_search_local()          # asynchronous: competes for the SMs while dem's search runs
        dem.uploadToGPU(dem.get_signalBufferHostPointer())
        res, E = dem._engine.search()
        v, o = dem._engine.peaks()
        out.append((E.copy(), v.copy(), o.copy()))
    for E, v, o in out[1:]:
        assert np.array_equal(out[0][0], E) and np.array_equal(out[0][1], v) and np.array_equal(out[0][2], o)


def test_stx_backend_matches_oracle():
    from pycusdr_b200.demodulator import STX
    conf = conf_variant("benchmark/bench_GMSK.json", blockSize=15)
    conf["GPU"]["UHF"]["peakThresholdScale"] = 4.5
    P = protocol_for(conf)
    dem = STX.Demodulator(conf, P, RADIO)
    orc = O.OracleDemodulator(conf, P, RADIO, backend="STX")
    sig, _ = S.bench_stream("GMSK", 15, seed=3)
    sig = sig.copy()
    sig[40000:40003] *= 60          # an interference burst that must get clipped and tagged
    got, ref = O.run_stream(dem, sig.copy()), O.run_stream(orc, sig.copy())
    for g, r in zip(got[1:], ref[1:]):      # chunk 0 holds the zero-filled overlap (see the stream test)
        np.testing.assert_array_equal(g["data"], r["data"])
    np.testing.assert_array_equal(dem.clippedPeakIPure, orc.clippedPeakIPure)


def _stx_pair(blockSize=15, scale=4.5):
    from pycusdr_b200.demodulator import STX
    conf = conf_variant("benchmark/bench_GMSK.json", blockSize=blockSize)
    conf["GPU"]["UHF"]["peakThresholdScale"] = scale
    P = protocol_for(conf)
    return STX.Demodulator(conf, P, RADIO, native_threshold=True), STX.Demodulator(conf, P, RADIO, native_threshold=False)


@pytest.mark.parametrize("blockSize", [12, 15, 18])
def test_device_threshold_is_numpy_exact_when_magnitudes_are(blockSize):
    """pcs_upload_thresholded (a19, dem_base:670-707) on samples whose magnitude is exact in any implementation (one
    component zero): clip levels (np.mean's pairwise float32 order), clipped indices and the clipped pinned buffer must be
    bit-identical to the NumPy statement of the reference."""
    dev, host = _stx_pair(blockSize)
    N = dev.Nfft
    rng = np.random.RandomState(blockSize)
    x = np.zeros(N, np.complex64)
    x.real[::2] = rng.randn(N // 2).astype(np.float32)
    x.imag[1::2] = rng.randn(N // 2).astype(np.float32)
    for pos, val in ((100, 40.0), (101, -35.0j), (150, 25.0), (N // 2, 1e3j), (N - 3, -77.0), (N - 2, 9.5j)):
        x[pos] = np.complex64(val)
    for d in (dev, host):
        d.get_signalBufferHostPointer()[:] = x
        d.uploadAndFindCarrier(d.get_signalBufferHostPointer())
    a, b = dev.get_signalBufferHostPointer(), host.get_signalBufferHostPointer()
    assert len(host.clippedPeakIPure) >= 5
    np.testing.assert_array_equal(dev.clippedPeakIPure, host.clippedPeakIPure)
    np.testing.assert_array_equal(dev.clippedPeakI, host.clippedPeakI)
    np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))
    # the second clip level is what NumPy computes from the (identical) clipped magnitudes of pass one
    mag = np.abs(x)
    t1 = np.float32(4.5) * np.mean(mag)
    assert dev.clipLevels[0] == t1
    # and the device copy of the chunk is the clipped one (what the demod stage then works on)
    a, b = dev.demodulate(), host.demodulate()
    for u, v in zip(a[:3], b[:3]):
        np.testing.assert_array_equal(u, v)


def test_device_threshold_matches_numpy_on_a_stream():
    """General complex samples: magnitudes differ in the last ulp between hypotf and np.abs, so the clip levels agree to
    float32 rounding; bursts far above the level give identical index lists, samples agree to 1e-6, bits are identical."""
    dev, host = _stx_pair()
    sig, _ = S.bench_stream("GMSK", 15, seed=3)
    sig = sig.copy()
    sig[40000:40003] *= 60
    sig[70000:70200:7] *= 45
    N, ovl = dev.Nfft, dev.sigOverlap
    ra, rb = dev.get_signalBufferHostPointer(), host.get_signalBufferHostPointer()
    ra[:] = 0
    rb[:] = 0
    clipped = 0
    for c in range(len(sig) // (N - ovl)):
        blk = sig[c * (N - ovl):(c + 1) * (N - ovl)]
        ra[ovl:] = blk
        rb[ovl:] = blk
        dev.uploadAndFindCarrier(ra)
        host.uploadAndFindCarrier(rb)
        np.testing.assert_array_equal(dev.clippedPeakIPure, host.clippedPeakIPure)
        np.testing.assert_allclose(ra.view(np.float32), rb.view(np.float32), rtol=1e-6, atol=1e-7)
        a, b = dev.demodulate(), host.demodulate()
        if c > 0:
            np.testing.assert_array_equal(a[0], b[0])
            # (the trust bytes are the low mantissa bytes of the float magnitudes, dem_base:1005-1007: they follow the
            #  1e-6 differences of the clipped samples and are compared in the exact-magnitude test above instead)
        clipped += len(dev.clippedPeakIPure)
        ra[:ovl] = ra[-ovl:]
        rb[:ovl] = rb[-ovl:]
    assert clipped > 30


def test_graph_replay_and_eager_launches_agree():
    """The per-chunk sequence replayed as a CUDA graph must give the very same bytes as launching it kernel by kernel."""
    conf = load_conf("benchmark/bench_GMSK.json")
    demA, _ = _demods(conf, use_graph=True)
    demB, _ = _demods(conf, use_graph=False)
    sig, _ = S.bench_stream("GMSK", 12, seed=6)
    a, b = O.run_stream(demA, sig), O.run_stream(demB, sig)
    assert len(a) > 5
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x["data"], y["data"])
        np.testing.assert_array_equal(x["trust"], y["trust"])
        np.testing.assert_equal(x["doppler"], y["doppler"])
        np.testing.assert_equal(x["doppler_std"], y["doppler_std"])
        np.testing.assert_equal(x["SNR"], y["SNR"])
    assert demA._engine.launch_count == demB._engine.launch_count


def test_snr_matches_oracle_on_a_stream():
    """computeSNR (dem_base:635-667) from the pruned spectrum bins vs the oracle's full-spectrum evaluation."""
    conf = load_conf("benchmark/bench_GMSK.json")
    dem, orc = _demods(conf)
    sig, _ = S.bench_stream("GMSK", 15, seed=8)
    for (fa, fb, _, _, _, _) in _run_both(dem, orc, sig):
        np.testing.assert_allclose(fa[3], fb[3], rtol=2e-4, atol=2e-4, equal_nan=True)
        np.testing.assert_allclose(fa[1], fb[1], rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("cfg,blockSize,sum_all", [("benchmark/bench_GMSK.json", 15, True), ("benchmark/bench_GMSK.json", 14, False),
                                                   ("CC11xx.json", 16, True), ("benchmark/bench_BPSK.json", 13, False)])
def test_parseval_variant_matches_surface_energies(cfg, blockSize, sum_all):
    """The labelled Parseval variant: same energies as the correlation-surface path without any inverse
    transform, hence the same Doppler estimate, shift and demodulated bits; it reports no peak."""
    from pycusdr_b200 import _native
    from pycusdr_b200.demodulator import UHF
    conf = conf_variant(cfg, blockSize=blockSize)
    P = protocol_for(conf)
    P.SUM_ALL_MASKS_PYTHON = sum_all
    demP = UHF.Demodulator(conf, P, RADIO, path=_native.PATH_PARSEVAL)
    demS = UHF.Demodulator(conf, P, RADIO)
    orc = O.OracleDemodulator(conf, P, RADIO)
    assert demP._engine.plan()["path"] == _native.PATH_PARSEVAL
    sps = conf["Radios"]["Rx"][RADIO]["samplesPerSym"]
    N = 2 ** blockSize
    for seed in (5, 6, 7):                          # several chunks: the second one on replays the CUDA graph
        x = _noise_chunk(N, seed, with_packet="GMSK" if sps == 16 else None, sps=sps)
        for d in (demP, demS):
            d.get_signalBufferHostPointer()[:] = x
        fp = demP.uploadAndFindCarrier(demP.get_signalBufferHostPointer())
        fs = demS.uploadAndFindCarrier(demS.get_signalBufferHostPointer())
        bp, bs = demP.demodulate(), demS.demodulate()
        Eo = O.search_energy_parseval(O.forward_fft(x), orc.masks, orc.doppCyperSymNorm, sum_all)
        assert rel_err(demP.last["E"], Eo) < 2e-5
        assert rel_err(demP.last["E"], demS.last["E"]) < 1e-4
        assert demP.last["shift"] == demS.last["shift"]
        assert fp[0] == pytest.approx(fs[0], abs=1e-2)
        np.testing.assert_allclose(fp[3], fs[3], rtol=1e-4, atol=1e-4, equal_nan=True)      # SNR: gathered vs pruned bins
        np.testing.assert_array_equal(bp[0], bs[0])
        assert demP.last["peak"] == (-1.0, -1, -1, -1)


@pytest.mark.parametrize("inflight,blk", [(1, 4096), (2, 10007), (3, 65536)])
def test_streaming_ingest_equals_the_chunk_loop(inflight, blk):
    """pcs_ingest_* (native ring + device-side overlap carry + chunks in flight) must reproduce, bit for bit, what the
    reference's chunk loop produces through uploadAndFindCarrier / demodulate on the same samples."""
    from pycusdr_b200.demodulator import UHF
    from pycusdr_b200.demodulator.stream import StreamDemodulator
    conf = load_conf("benchmark/bench_GMSK.json")
    P = protocol_for(conf)
    sig, _ = S.bench_stream("GMSK", 11, seed=9)
    ref = O.run_stream(UHF.Demodulator(conf, P, RADIO), sig)
    sd = StreamDemodulator(conf, P, RADIO, inflight=inflight)
    got = []
    for a in range(0, len(sig), blk):
        got += sd.push(sig[a:a + blk])
    got += sd.flush()
    assert len(got) == len(ref) > 5
    for g, r in zip(got, ref):
        np.testing.assert_array_equal(g["data"], r["data"])
        np.testing.assert_array_equal(g["trust"], r["trust"])
        np.testing.assert_equal(g["doppler"], r["doppler"])
        np.testing.assert_equal(g["SNR"], r["SNR"])
        assert g["spSymEst"] == r["spSymEst"]


@pytest.mark.parametrize("fused", [True, False])
def test_chunks_from_a_registered_sample_ring_equal_the_fill_loop(fused):
    """Extension of the class contract (Demodulator.registerHostMemory, pcs_host_register / pcs_set_host_source): chunks
    passed as overlapping windows of a page-locked sample ring go to the GPU from there, without the caller's fill of the one
    pinned chunk buffer (demodulator_process.py:287) -- every output must equal what the unchanged caller loop produces."""
    from pycusdr_b200 import sharded
    from pycusdr_b200.demodulator import UHF
    conf = load_conf("benchmark/bench_GMSK.json")
    P = protocol_for(conf)
    sig, _ = S.bench_stream("GMSK", 11, seed=19)
    ref = O.run_stream(UHF.Demodulator(conf, P, RADIO, fused=fused), sig)
    dem = UHF.Demodulator(conf, P, RADIO, fused=fused)
    N, ovl = dem.Nfft, dem.sigOverlap
    step = N - ovl
    ring = np.concatenate((np.zeros(ovl, np.complex64), np.ascontiguousarray(sig, dtype=np.complex64)))
    before = ring.copy()
    reg = dem.registerHostMemory(ring)
    own = dem.get_signalBufferHostPointer()
    own[:] = np.complex64(7 + 7j)             # must not be what gets transformed
    assert len(ref) > 5
    for c, r in enumerate(ref):
        freq, sdev, clipped, snr = dem.uploadAndFindCarrier(ring[c * step:c * step + N])
        bits, centres, trust, sp = dem.demodulate()
        np.testing.assert_array_equal(bits, r["data"])
        np.testing.assert_array_equal(trust, r["trust"])
        np.testing.assert_equal(freq, r["doppler"])
        np.testing.assert_equal(snr, r["SNR"])
        assert sp == r["spSymEst"]
    np.testing.assert_array_equal(ring, before)            # read in place, never written
    assert np.all(own == np.complex64(7 + 7j))
    # the streaming engine takes the same windows (PCS_SRC_HOST with a caller pointer): identical bit stream
    dem2 = UHF.Demodulator(conf, P, RADIO)
    sh = sharded.ShardedStream(dem2._engine, 0, 1, lambda o: [o], lag=2)
    bs = sharded.ShardedBitStream(sh, dem2._stitch, 0, 1, None, None)
    for c in range(len(ref)):
        bs.submit(ring[c * step:c * step + N], sharded.SRC_HOST)
    bs.finish()
    for c, r in enumerate(ref):
        np.testing.assert_array_equal(bs.bits[c][0], r["data"])
        np.testing.assert_array_equal(bs.bits[c][2], r["trust"])
    reg.close()
    # a window of memory that is not page-locked takes the copy path (and still gives the same answer) ...
    dem.uploadAndFindCarrier(ring[0:N].copy())
    # ... and the native call refuses a pageable source instead of silently staging it
    from pycusdr_b200 import _native
    with pytest.raises(_native.NativeError):
        dem._engine.set_host_source(before.__array_interface__["data"][0])
