"""Host-side mirror of the reference's ``Demodulator`` (pyCuSDR/demodulator/demodulator_base.py).

Same constructor, same public methods, same return types; the device work is done by the
sm_100a kernels behind ``include/pycusdr_b200.h``.  What stays in Python here is exactly what the
reference keeps on the host: configuration (dem_base:84-174), the Hz interpolation and SNR
arithmetic (:610-623, :635-667), bit extraction (:1012-1051), chunk stitching (:863-988), clipped
peak tagging (:817-837), the STX input thresholding (:670-707) and the output casts (:859).

Differences from the reference that a caller can observe (DESIGN.md "quirks"):
  * search and demodulation of a chunk are enqueued together in ``uploadAndFindCarrier`` (one
    synchronisation per chunk instead of five), and with the native bit post-processing the whole
    chunk is ONE native call (``pcs_chunk_to_bits``); ``demodulate`` then only hands the bits over.
    ``one_call=False`` keeps the post-processing in ``demodulate``; ``fused=False`` gives the
    reference's two-step device schedule.  All three produce identical outputs.
  * ``E`` is reduced in a fixed order, so results are bit-reproducible run to run (the reference's
    float atomics are not, kern:463,474).
  * there is no ``extractBitsOld`` (undefined in the reference too, dem_base:1017).
"""
import logging

import numpy as np
import scipy.constants

from .. import LOG_NAME
from .. import _native

log = logging.getLogger(LOG_NAME + "." + __name__)

# Defaults for the symbol overlap test (dem_base:19-22)
SYMBOL_CHECK_OVERLAP_OFFSET = 20
SYMBOL_CHECK_ERROR_THRESHOLD = 1000
SYMBOL_CHECK_MATCH_NUM_ERRORS_ALLOWED = 10
SYMBOL_MISMATCHVAL = 0      # dem_base:26
TRUSTTYPE = np.int8         # __global__.py:24
SNR_WINDOW = 5              # dem_base:620


class Demodulator:

    def __init__(self, conf, protocol, radioName, fused=True, path=_native.PATH_AUTO, log2_block=0, use_graph=True,
                 native_post=True, groups_per_cta=0, xb_smem=False, search_form=0, items_per_cta=0, one_call=True,
                 warps20=False, native_threshold=False):
        self.protocol = protocol
        self.radioName = radioName
        self.confRadio = confRadio = conf["Radios"]["Rx"][radioName]
        self.confGPU = confGPU = conf["GPU"][confRadio["CUDA_settings"]]
        self.fused = fused
        self.one_call = one_call
        # STX backend: input clipping on the device (opt-in: agrees with the NumPy statement to 1e-6, not bit for bit,
        # because np.abs(complex64) is a host-dependent SIMD approximation; see STX.py)
        self.native_threshold = native_threshold

        # chunk geometry (dem_base:89-93)
        self.sigLen = 2 ** confGPU["blockSize"]
        self.sigOverlap = 2 ** confGPU["overlap"]
        self.sigOverlapWin = int(self.sigOverlap / 2)
        self.clippedPeakSpan = confGPU["clippedPeakSpan"]
        self.peakThresholdScale = confGPU["peakThresholdScale"]
        self.disablePeakThresholding = confRadio.get("disablePeakThresholding", False)
        # symbol overlap check (dem_base:97-99)
        self.overlapOffset = confGPU.get("symbol_check_overlap_offset", SYMBOL_CHECK_OVERLAP_OFFSET)
        self.symbol_check_error_threshold = confGPU.get("symbol_check_error_threshold", SYMBOL_CHECK_ERROR_THRESHOLD)
        self.symbol_check_match_threshold = self.overlapOffset - confGPU.get(
            "symbol_check_match_num_errors_allowed", SYMBOL_CHECK_MATCH_NUM_ERRORS_ALLOWED)
        log.info(f"[{radioName}]: symbol_check_overlap_offset {self.overlapOffset}, symbol_check_error_threshold "
                 f"{self.symbol_check_error_threshold}, symbol_check_match_threshold {self.symbol_check_match_threshold}")
        # radio (dem_base:103-113)
        self.spsym = spsym = confRadio["samplesPerSym"]
        self.spsymMin = int(spsym / 2)
        self.baudRate = confRadio["baud"]
        self.sampleRate = self.baudRate * self.spsym
        self.voteWeight = confRadio.get("voteWeight", 1)
        self.Nfft = int(self.sigLen)
        self.windowWidth = confGPU["bitWindowWidth"]
        self.windowWidthOffset = int(self.windowWidth / 2)
        self.CODE_SEARCH_MASK_OFFSET = 0                                    # dem_base:120
        self.SUM_ALL_MASKS_PYTHON = bool(getattr(protocol, "SUM_ALL_MASKS_PYTHON", False))
        log.info(f"[{radioName}]: Sum masks prior to Doppler search {self.SUM_ALL_MASKS_PYTHON}")

        # Doppler grid (dem_base:130-165)
        self.num_dopplers = confRadio["doppCarrierSteps"]
        self.centreFreqOffset = confRadio["frequencyOffset_Hz"]
        Fc = confRadio["frequency_Hz"] - self.centreFreqOffset
        self.doppOffset = self.centreFreqOffset / self.baudRate / self.spsym
        self.doppOffsetIdx = np.int32(self.doppOffset * self.Nfft)
        if self.doppOffsetIdx < 0:
            self.doppOffsetIdx += self.Nfft
        doppMaxNorm = conf["Radios"]["rangeRateMax"] * Fc / scipy.constants.speed_of_light / self.sampleRate
        grid = np.linspace(self.doppOffset - doppMaxNorm, self.doppOffset + doppMaxNorm, self.num_dopplers)
        noiseOfftuneHz = confRadio.get("noise_measure_offset_Hz", False)
        if noiseOfftuneHz:
            grid = np.concatenate((np.array([noiseOfftuneHz / self.baudRate / self.spsym]), grid))
        self.doppIdxNorm = grid
        self.doppIdxArrayLen = len(grid)
        self.doppIdxArrayOffset = self.doppIdxArrayLen - self.num_dopplers
        self.doppHzLUT = grid * self.spsym * self.baudRate
        self.doppCyperSymNorm = np.round(grid * self.Nfft).astype(np.int32)
        self.doppCyperSymNorm[self.doppCyperSymNorm < 0] += self.Nfft
        log.info("[{}]: Fc {:.0f} Doppler scanning range {:.0f} to {:.0f} Hz of Fc".format(
            radioName, Fc, self.doppHzLUT[0], self.doppHzLUT[-1]))

        # filters and LUTs from the protocol plugin (dem_base:194-211, 246-257)
        try:
            self.num_masks, masks = protocol.get_filter(self.Nfft, self.spsym, confGPU["xcorrMaskSize"])
            if masks.shape != (self.num_masks, self.Nfft):
                raise ValueError("Masks provided by protocol {} expected to be of dimensions {}, got dimensions {}".format(
                    protocol.name, (self.num_masks, self.Nfft), masks.shape))
            if not isinstance(masks[0, 0], np.complex64):
                raise TypeError("Datatype of masks {}, expected {}".format(type(masks[0, 0]), np.complex64))
            if self.num_masks > 32:
                log.warning("[{}]: more than 32 masks is not supported at this time".format(radioName))
        except Exception:
            log.error("[{}]: Exception occured in protocol {} while preparing filters".format(radioName, protocol.name))
            raise
        try:
            self.bitLUT, self.symbolLUT = protocol.get_symbolLUT2(confGPU["xcorrMaskSize"])
        except Exception:
            log.error("[{}]: Exception occured in protocol {} while preparing symbol lookup table".format(
                radioName, protocol.name))
            raise

        self.symsTolLow = 0.9 * self.spsym                                  # dem_base:508-512
        self.symsTolHigh = 1.1 * self.spsym
        self.codeRateAndPhaseOffsetLow = int(self.Nfft / self.symsTolLow)
        self.codeRateAndPhaseOffsetHigh = int(self.Nfft / self.symsTolHigh)

        # device side: raises (no CPU fallback) if the library or a GPU is missing
        self._engine_kwargs = dict(
            device=confGPU["CUDA"]["device"], nfft=self.Nfft, num_dopplers=self.num_dopplers,
            element_offset=self.doppIdxArrayOffset, shifts=self.doppCyperSymNorm, masks=masks,
            window_width=self.windowWidth, sum_all_masks=self.SUM_ALL_MASKS_PYTHON,
            code_search_mask_offset=self.CODE_SEARCH_MASK_OFFSET, samples_per_sym=self.spsym, path=path,
            log2_block=log2_block, snr_window=SNR_WINDOW, use_graph=use_graph, groups_per_cta=groups_per_cta, xb_smem=xb_smem,
            search_form=search_form, items_per_cta=items_per_cta, warps20=warps20)
        self._engine = _native.Engine(**self._engine_kwargs)
        self.GPU_bufSignalTime_cpu_handle = self._engine.host_buffer
        # bit extraction + chunk stitching + trust tagging in one native call (the NumPy methods below stay as the
        # readable mirror of dem_base:863-1051 and are what ``native_post=False`` runs)
        self._stitch = None
        if native_post and (self.bitLUT is not None or len(np.shape(self.symbolLUT)) == 3):
            self._stitch = _native.Stitcher(
                nfft=self.Nfft, overlap=self.sigOverlap, overlap_offset=self.overlapOffset,
                error_threshold=self.symbol_check_error_threshold, match_threshold=self.symbol_check_match_threshold,
                bit_lut=self.bitLUT, symbol_lut=self.symbolLUT)

        # cross-call state
        self.clippedPeakIPure = []
        self.clippedPeakI = []
        self.poswinP = []
        self.posSymEnd = None
        self.dopplerIdxlast = 0
        self._pending = None        # device results of the chunk currently being processed
        self._bits_ready = None     # ... and its finished bits when the whole chunk ran in one native call
        self.last = {}              # inspection: raw per-chunk device outputs
        log.info("[{}]: Initialization done ({})".format(radioName, self._engine.plan()))

    # ------------------------------------------------------------------------------------------
    def __del__(self):
        eng = getattr(self, "_engine", None)
        if eng is not None:
            self.GPU_bufSignalTime_cpu_handle = None
            eng.close()

    def get_signalBufferHostPointer(self):
        """Pinned complex64[Nfft] the caller fills in place (dem_base:1055-1060)."""
        return self.GPU_bufSignalTime_cpu_handle

    def _as_chunk_buffer(self, samples, direct=False):
        """The reference transforms the pinned buffer whatever ``samples`` is (dem_base:557); a caller
        that passes another array gets it copied in, which is what it meant.  ``direct``: a complex64[Nfft] view of memory
        the caller page-locked with ``registerHostMemory`` is NOT copied -- the next upload reads it in place."""
        buf = self.GPU_bufSignalTime_cpu_handle
        if samples is buf:
            return buf
        if direct:
            ptr = _native.registered_ptr(samples, 8 * self.Nfft) if getattr(samples, "dtype", None) == np.complex64 else None
            if ptr is not None:
                self._engine.set_host_source(ptr)
                return samples
        if not (isinstance(samples, np.ndarray) and np.shares_memory(samples, buf)):
            buf[:] = samples
        return buf

    @staticmethod
    def registerHostMemory(arr):
        """Extension: page-lock the caller's own sample memory (e.g. the ring a receiver thread fills; consecutive chunks are
        overlapping windows of it).  ``uploadAndFindCarrier(view)`` with a complex64[Nfft] view of it then copies to the GPU
        straight from there: the fill of the pinned chunk buffer (demodulator_process.py:287, 8 * Nfft bytes of host memcpy
        per chunk) and the overlap carry (:337) disappear.  Returns the registration (``close()`` releases the page lock)."""
        return _native.host_register(arr)

    # -- a4 --------------------------------------------------------------------------------------
    def uploadToGPU(self, samples, direct=False):
        self._as_chunk_buffer(samples, direct)
        self._engine.upload()
        self._pending = None
        self._bits_ready = None

    def thresholdInput(self, samples):
        self.__thresholdInput(samples)

    def thresholdAndUpload(self, samples):
        """__thresholdInput + uploadToGPU with the clipping done on the device (``pcs_upload_thresholded``); the pinned
        buffer ends up clipped in place like the reference leaves it (dem_base:670-707, 548-558) -- to 1e-6 relative in
        the clipped samples and the two clip levels, exactly when the magnitudes are exact in any implementation.  The
        level is float32(scale) * float32 mean, which is what ``peakThresholdScale * np.mean(mag)`` evaluates to under
        NumPy >= 2 (a NumPy 1.x host rounds the float64 product once instead)."""
        self._as_chunk_buffer(samples)
        over, self.clipLevels = self._engine.upload_thresholded(np.float32(self.peakThresholdScale))
        self._pending = None
        self._bits_ready = None
        self.clippedPeakIPure = over
        self.peakMinGap = 100
        self.clippedPeakI = _native.fill_gaps(over, self.peakMinGap, self.Nfft) if len(over) else over.copy()

    def uploadAndFindUHF(self, samples):
        samples = self._as_chunk_buffer(samples)
        self.__thresholdInput(samples)
        self.uploadToGPU(samples)
        return self.findUHF(samples)

    def chunkToBits(self, samples):
        """uploadToGPU + findUHF + the whole of demodulate() in ONE native call (``pcs_chunk_to_bits``): the fused
        schedule with the symbol tables handed to the stitcher inside the library.  Returns what findUHF returns; the
        bits are parked for the demodulate() call that follows (dem_base:548-632, 765-859)."""
        self._as_chunk_buffer(samples, direct=True)
        (res, E, sym, centre, mag, means, bits, centres8, trust8, err) = self._engine.chunk_to_bits(
            self._stitch, self.doppCyperSymNorm, self.clippedPeakIPure)
        self._pending = None
        self._bits_ready = (res, sym, centre, mag, bits, centres8, trust8, err)
        return self._finish_search(res, E, means=means)

    # -- a5 --------------------------------------------------------------------------------------
    def findUHF(self, samples=None):
        """Doppler search (dem_base:567-632). Returns (freqOffset, sdev_Hz, clippedPeakIPure, SNR)."""
        eng = self._engine
        if self.fused:
            res, E, sym, centre, mag = eng.process()
            self._pending = (res, sym.copy(), centre.copy(), mag.copy())
        else:
            res, E = eng.search()
            self._pending = None
        return self._finish_search(res, E)

    def findUHFRates(self, samples, rates_hz_per_s):
        """Extension: Doppler-RATE search dimension.  The reference prepares ``complexHeterodyne``
        (kern:755-778, dem_base:388) and never calls it; here every rate hypothesis de-chirps the uploaded chunk with that
        kernel's statement (``pcs_heterodyne``, a = -pi r / fs^2) and runs the ordinary Doppler search on it; the hypothesis
        with the largest mask-summed energy wins (first one on ties), and the chunk stays de-chirped with it for the
        ``demodulate()`` that follows.  Returns (best rate, metric per rate) + what ``findUHF`` returns."""
        self.uploadToGPU(samples)
        eng = self._engine
        fs = float(self.sampleRate)
        coeff = [np.float32(-np.pi * float(r) / fs ** 2) for r in rates_hz_per_s]
        metrics = []
        for a in coeff:
            eng.heterodyne(a)
            res, E = eng.search()
            metrics.append(float(np.max(np.sum(E[self.doppIdxArrayOffset:], axis=1, dtype=np.float64))))
        best = int(np.argmax(metrics))
        eng.heterodyne(coeff[best])
        self.dopplerRate = float(rates_hz_per_s[best])
        return (self.dopplerRate, np.array(metrics)) + tuple(self.findUHF())

    def _finish_search(self, res, E, wins=None, means=None):
        """Host half of __findUHF (dem_base:604-632) on the device results of one chunk."""
        self.last = {"E": E.copy(), "res": np.array([res.best_idx, res.metric_db], dtype=np.float32),
                     "peak": (res.peak_val, res.peak_bin, res.peak_mask, res.peak_offset)}
        if res.status != 0:          # NaN estimate: the reference's ValueError branch (dem_base:625-630)
            log.error("Error occurred during find_UHF -- skipping block. Message: cannot convert float NaN to integer")
            self.dopplerIdxlast = 0
            return 0., 0., self.clippedPeakIPure, 0.
        lowIdx, highIdx = res.low_idx, res.high_idx
        frac = np.float64(res.best_idx) % 1
        lowVal, highVal = self.doppHzLUT[lowIdx], self.doppHzLUT[highIdx]
        bestDopplerScaled = lowVal + (highVal - lowVal) * frac
        self.dopplerIdxlast = np.int32(res.shift)
        SNR = self.computeSNR(lowIdx, highIdx, SNR_WINDOW, res, wins, means)
        freqOffset = bestDopplerScaled - self.centreFreqOffset
        sdev_Hz = np.float64(res.metric_db) / self.Nfft * self.sampleRate
        return freqOffset, sdev_Hz, self.clippedPeakIPure, SNR

    # -- a10 -------------------------------------------------------------------------------------
    def computeSNR(self, doppMatchLow, doppMatchHigh, windowWidth, res=None, given=None, means=None):
        """SNR from the chunk spectrum around the found bins vs the same window half a band away
        (dem_base:635-667), evaluated with the reference's slicing rules on windows the device gathered."""
        N = self.Nfft
        lo = int(self.doppCyperSymNorm[doppMatchLow])
        hi = int(self.doppCyperSymNorm[doppMatchHigh])
        nlo, nhi = (lo + N // 2) % N, (hi + N // 2) % N
        wins = None
        if means is not None:        # the library already averaged the two windows (common geometry, pcs_snr_means)
            with np.errstate(all="ignore"):
                return np.float64(20) * np.log10(np.float64(means[0] / means[1]) - 1)
        if res is not None and res.sig_len > 0 and windowWidth == SNR_WINDOW:
            sig, noise = given if given is not None else self._engine.snr_windows(res)
            # common case: neither window touches the ends of the spectrum, so the reference's slices
            # X[a-w : b+w] are exactly the two windows the device gathered (dem_base:657-661)
            w = windowWidth
            if (lo <= hi and nlo <= nhi and lo - w >= 0 and nlo - w >= 0 and hi + w <= N and nhi + w <= N
                    and res.sig_start == lo - w and res.noise_start == nlo - w
                    and len(sig) == hi - lo + 2 * w and len(noise) == nhi - nlo + 2 * w):
                with np.errstate(all="ignore"):
                    ratio = _native.mean_abs_c64(sig) / _native.mean_abs_c64(noise)     # float32 / float32
                    return np.float64(20) * np.log10(np.float64(ratio) - 1)
            wins = ((res.sig_start, sig), (res.noise_start, noise))
        full = [None]

        def gather(ranges, which):
            idx = np.concatenate([np.arange(r.start, r.stop, r.step) for r in ranges]) if ranges else np.array([], int)
            if wins is not None:
                start, win = wins[which]
                pos = (idx - start) % N
                if idx.size == 0 or pos.max() < len(win):
                    return win[pos]
            if full[0] is None:
                full[0] = self._engine.spectrum()
            return full[0][idx]

        def window_mean(a, b, which):
            if a > b:        # the window wraps around bin 0 (dem_base:653-654)
                ranges = [range(N)[a - windowWidth:], range(N)[:b + windowWidth]]
            else:
                ranges = [range(N)[a - windowWidth:b + windowWidth]]
            return np.mean(np.abs(gather(ranges, which)))

        with np.errstate(all="ignore"):
            sigPwr = window_mean(lo, hi, 0)
            noisePwr = window_mean(nlo, nhi, 1)
            ratio = np.float32(sigPwr) / np.float32(noisePwr)
            return np.float64(20) * np.log10(np.float64(ratio) - 1)

    # -- a19 -------------------------------------------------------------------------------------
    def __thresholdInput(self, samples):
        """Two-pass clipping of strong bursts, in place (dem_base:670-707)."""
        mag = np.abs(samples)
        for rnd in range(2):
            thresh = self.peakThresholdScale * np.mean(mag)
            over = np.flatnonzero(mag > thresh)
            samples[over] = thresh * (samples[over] / mag[over])
            if rnd == 0:
                mag[over] = np.abs(samples[over])
        self._set_clipped(over)

    def _set_clipped(self, over):
        """clippedPeakIPure / clippedPeakI from the indices the second pass clipped (dem_base:683-705)."""
        self.clippedPeakIPure = over
        if len(over) > 0:
            self.peakMinGap = 100
            # runs of clipped samples separated by fewer than peakMinGap samples are merged
            step = np.diff(over)
            marks = np.zeros(self.Nfft, dtype=np.int8)
            marks[over] = 1
            for k in np.flatnonzero((step > 1) & (step < self.peakMinGap)):
                marks[over[k]:over[k] + step[k]] = 1
            self.clippedPeakI = np.flatnonzero(marks == 1)
        else:
            self.clippedPeakI = over.copy()

    # -- a11 -------------------------------------------------------------------------------------
    def demodulateUHF(self):
        return self.__demodulate()

    def demodulateSTX(self):
        self.dopplerIdxlast = self.doppOffsetIdx          # dem_base:760
        self._pending = None
        self._bits_ready = None
        return self.__demodulate()

    def __demodulate(self):
        if self._bits_ready is not None:
            res, idxSymbol, centres, magnitudes, bits, centres8, trust8, err = self._bits_ready
            self._bits_ready = None
            spSym = np.float64(res.sp_sym)
            self.last.update(shift=int(res.demod_shift), timing=np.array(res.timing[:], dtype=np.float32), spSym=spSym,
                             codeOffset=np.float64(res.code_offset), sym=idxSymbol.copy(), centres=centres.copy(),
                             mag=magnitudes.copy())
            if err is not None:
                raise err
            return bits, centres8, trust8, spSym
        if self._pending is not None:
            res, idxSymbol, centres, magnitudes = self._pending
            self._pending = None
        else:
            res, idxSymbol, centres, magnitudes = self._engine.demod(int(self.dopplerIdxlast))
        spSym = np.float64(res.sp_sym)
        # the reference reads the float magnitudes back through an int8-sized buffer (dem_base:472,1005-1007)
        trustSymbol = np.ascontiguousarray(magnitudes, dtype=np.float32).view(TRUSTTYPE)[:len(idxSymbol)].copy()
        self.last.update(shift=int(res.demod_shift), timing=np.array(res.timing[:], dtype=np.float32), spSym=spSym,
                         codeOffset=np.float64(res.code_offset), sym=idxSymbol, centres=centres, mag=magnitudes)
        if self._stitch is not None:
            bitsU8, centresU8, trustU8 = self._stitch(idxSymbol, centres, magnitudes, self.clippedPeakIPure, spSym)
            return bitsU8, centresU8, trustU8, spSym
        dataBits, symError = self.extractBits(centres, idxSymbol)
        centresWin, dataBitsWin, trustSymbolWin, _ = self.checkSymbolOverlap(
            len(symError), centres, idxSymbol, dataBits, trustSymbol)
        # tag symbols near clipped input peaks in the trust (dem_base:831-837)
        if len(self.clippedPeakIPure) > 0:
            near = np.zeros(self.Nfft, dtype=bool)
            span = 2 * int(np.ceil(spSym))
            for cp in self.clippedPeakIPure:
                near[cp - span:cp + span + 1] = True
            trustSymbolWin = trustSymbolWin.copy()
            trustSymbolWin[near[centresWin]] = -2
        return (dataBitsWin.astype(np.uint8), centresWin.astype(np.uint8), trustSymbolWin.astype(np.uint8), spSym)

    # -- a16 -------------------------------------------------------------------------------------
    def extractBits(self, centres, symbols):
        if self.bitLUT is None:
            if len(np.shape(self.symbolLUT)) == 3:
                return self.extractBitsNRZs(centres, symbols)
            raise NotImplementedError("extractBitsOld is not defined by the reference either (dem_base:1017)")
        return np.asarray(self.bitLUT)[symbols], []

    def extractBitsNRZs(self, centresCoherent, symbols):
        """NRZ-S decisions from symbol transitions (dem_base:1026-1051)."""
        lut = np.asarray(self.symbolLUT)
        cur, nxt = symbols[:-1], symbols[1:, None]
        ones = (lut[cur, 0, :] == nxt).any(axis=1)
        zeros = (lut[cur, 1, :] == nxt).any(axis=1)
        symError = np.flatnonzero(~(ones | zeros)).tolist()
        ones[symError] = int(SYMBOL_MISMATCHVAL)
        return ones, symError

    # -- a17 -------------------------------------------------------------------------------------
    def checkSymbolOverlap(self, noError, centres, idxSymbol, dataBits, trustSymbol):
        """Cut the chunk to [overlap/2, Nfft-overlap/2] and realign by +-1 symbol against the previous
        chunk's tail when the shifted comparison matches better (dem_base:863-988)."""
        oo = self.overlapOffset
        startOverlap = np.where(centres >= self.sigOverlapWin)[0][0]
        endOverlap = np.where(centres > (self.Nfft - self.sigOverlapWin))[0][0]
        win, pre = dataBits[startOverlap:endOverlap], dataBits[:startOverlap]

        def matches(a, b):
            a, b = np.asarray(a), np.asarray(b)
            return None if a.shape != b.shape else int(np.sum(a == b))

        if noError <= self.symbol_check_error_threshold and len(self.poswinP) > 0:
            P, Eend = self.poswinP, self.posSymEnd
            pairs = {
                "pre": (P[:oo], win[:oo]), "pos": (Eend[-oo:], pre[-oo:]),
                "earlyPre": (P[:oo], win[1:oo + 1]), "earlyPos": (Eend[-oo - 1:-1], pre[-oo:]),
                "latePre": (P[1:oo + 1], win[0:oo]), "latePos": (Eend[-oo:], pre[-oo - 1:-1]),
            }
            n = {k: matches(*v) for k, v in pairs.items()}
            full = lambda k: n[k] is not None and n[k] == len(pairs[k][0])   # noqa: E731  (np.all)
            if not (full("pre") or full("pos")):
                c = {k: (0 if v is None else v) for k, v in n.items()}     # length mismatch compares unequal
                maxPre = max(c["pre"], c["earlyPre"], c["latePre"])
                maxPos = max(c["pos"], c["earlyPos"], c["latePos"])
                thr = self.symbol_check_match_threshold
                if thr < c["earlyPre"] and c["earlyPre"] == maxPre:
                    if thr < c["earlyPos"] and c["earlyPos"] == maxPos:
                        startOverlap += 1                                    # drop the first bit
                elif thr < c["latePre"] and c["latePre"] == maxPre:
                    if thr < c["latePos"] and c["latePos"] == maxPos:
                        startOverlap -= 1                                    # re-insert the last pre-window bit
        dataBitsWin = dataBits[startOverlap:endOverlap]
        self.poswinP = dataBits[endOverlap:]
        self.posSymEnd = dataBitsWin[-oo - 1:]
        return (centres[startOverlap:endOverlap], dataBitsWin, trustSymbol[startOverlap:endOverlap],
                dataBitsWin)
