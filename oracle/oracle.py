"""CPU oracle for the pyCuSDR demodulator hot path -- TEST INFRASTRUCTURE ONLY.

This file is a NumPy/SciPy restatement of the reference's per-chunk algorithm (matched-filter
Doppler search, Doppler estimate, symbol-timing recovery, symbol decisions, bit extraction and
chunk stitching).  It exists to check the CUDA path in ``pycusdr_b200``; it is imported only by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs and is never on
the product path.

Parity status: the reference ships no test, golden vector or known-answer fixture for this
path, so the oracle is pinned as follows instead:
  * host-side functions (``extract_bits``, ``extract_bits_nrzs``, ``check_symbol_overlap``,
    ``threshold_input``, ``compute_snr``, the protocol filter banks and the signal
    generators) against the reference's own Python code imported from ``/root/reference``
    (``oracle/make_golden.py`` -> ``tests/golden/*.npz``, checked by ``tests/test_oracle_golden.py``);
  * device-side steps (the eight kernels of SURVEY.md 2.2(i)) against the reference's own
    ``cuda_kernels.cu`` compiled unmodified for sm_100a and run on a B200 through the
    driver API (``oracle/ref_gpu/``; outputs committed as ``tests/golden/refgpu_*.npz``).
Where a fixture of the second kind is absent the corresponding function is "parity unpinned"
and DESIGN.md says so.

Every function cites the reference lines it follows.  ``dem_base`` =
``pyCuSDR/demodulator/demodulator_base.py``, ``kern`` = ``pyCuSDR/demodulator/cuda_kernels.cu``.

Arithmetic: complex64 / float32 like the reference's device code, float64 where the
reference's host code (NumPy < 1.24 scalar promotion) computes in float64.  Fused
multiply-adds that nvcc contracts in the reference kernels (checked in the sm_100a SASS:
``findCentres`` FFMA x*spSym-3, ``findDopplerEst`` FFMA i0*v0 + (i1*v1)) are emulated through
float64, which is exact for float32 operands of this magnitude.
"""
import math

import numpy as np
import scipy.constants
import scipy.fft

F32 = np.float32
SYMBOL_CHECK_OVERLAP_OFFSET = 20            # dem_base:20
SYMBOL_CHECK_ERROR_THRESHOLD = 1000         # dem_base:21
SYMBOL_CHECK_MATCH_NUM_ERRORS_ALLOWED = 10  # dem_base:22
SYMBOL_MISMATCHVAL = 0                      # dem_base:26
E_SCALE = F32(262144.0)                     # kern:442 (2^18, independent of Nfft)


def _fma32(a, b, c):
    """float32 fma(a, b, c) for float32 inputs (product exact in float64)."""
    return F32(np.float64(a) * np.float64(b) + np.float64(c))


def abs2(z):
    """re^2 + im^2 in float32 (kern:1022-1026)."""
    z = np.asarray(z)
    return (z.real * z.real + z.imag * z.imag).astype(F32, copy=False)


def abs2_fma(z, order="xy"):
    """abs2 as nvcc contracts the reference's ComplexAbsSquared (kern:1022-1026) on sm_100a: one FMUL and one
    FFMA, i.e. fma(x, x, round(y*y)) ("xy") or fma(y, y, round(x*x)) ("yx").  Emulated through float64 (the
    product of two float32 is exact there).  Which operand order the compiler picked is established by
    tests/test_oracle_refgpu.py against the reference kernels' own output."""
    z = np.asarray(z)
    a, b = (z.real, z.imag) if order == "xy" else (z.imag, z.real)
    bb = (b.astype(F32) * b.astype(F32)).astype(F32)
    return (a.astype(np.float64) * a.astype(np.float64) + bb.astype(np.float64)).astype(F32)


# --------------------------------------------------------------------------------------
# a1: Doppler grid (dem_base:130-165)
# --------------------------------------------------------------------------------------
def doppler_grid(conf, radioName, Nfft):
    confRadio = conf["Radios"]["Rx"][radioName]
    spsym = confRadio["samplesPerSym"]
    baud = confRadio["baud"]
    fs = baud * spsym
    num_dopplers = confRadio["doppCarrierSteps"]
    centreFreqOffset = confRadio["frequencyOffset_Hz"]
    Fc = confRadio["frequency_Hz"] - centreFreqOffset
    doppOffset = centreFreqOffset / baud / spsym
    doppOffsetIdx = np.int32(doppOffset * Nfft)
    if doppOffsetIdx < 0:
        doppOffsetIdx += Nfft
    doppMaxNorm = conf["Radios"]["rangeRateMax"] * Fc / scipy.constants.speed_of_light / fs
    lo, hi = doppOffset - doppMaxNorm, doppOffset + doppMaxNorm
    noiseHz = confRadio.get("noise_measure_offset_Hz", False)
    grid = np.linspace(lo, hi, num_dopplers)
    if noiseHz:
        grid = np.concatenate((np.array([noiseHz / baud / spsym]), grid))
    shifts = np.round(grid * Nfft).astype(np.int32)
    shifts[shifts < 0] += Nfft
    return {
        "doppIdxNorm": grid,
        "shifts": shifts,
        "doppHzLUT": grid * spsym * baud,
        "num_dopplers": num_dopplers,
        "element_offset": len(grid) - num_dopplers,
        "doppOffsetIdx": doppOffsetIdx,
        "centreFreqOffset": centreFreqOffset,
        "fs": fs,
    }


# --------------------------------------------------------------------------------------
# a4: forward FFT (dem_base:557); a6-a8: surface and energy (kern:339-373, dem_base:578-591,
# kern:421-480)
# --------------------------------------------------------------------------------------
def forward_fft(x, dtype=np.complex64):
    return scipy.fft.fft(np.asarray(x, dtype=dtype))


def surface_rows(X, masks, shift):
    """y[m, :] = IFFT_unnormalised(X[(k+shift) % N] * Mk[m, k])  (kern:174-185 / :339-373)."""
    Y = np.roll(X, -int(shift))[None, :] * masks
    return scipy.fft.ifft(Y, axis=1, norm="forward")


def search_energy(X, masks, shifts, sum_all_masks, want_peaks=False, workers=1):
    """E[d, m] = sum_n abs2(y[d,m,n]) / 2^18 (+ optional per-(d,m) peak and its offset).

    SUM mode folds all masks into column 0 (kern:453-464); other columns stay 0 (they are zeroed
    by setArrayToZeros, kern:853-857, dem_base:571).
    """
    D, (M, N) = len(shifts), masks.shape
    E = np.zeros((D, M), dtype=F32)
    peak_val = np.zeros((D, M), dtype=F32)
    peak_off = np.zeros((D, M), dtype=np.int32)
    for d in range(D):
        Y = np.roll(X, -int(shifts[d]))[None, :] * masks
        y = scipy.fft.ifft(Y, axis=1, norm="forward", workers=workers)
        p = abs2(y)
        E[d] = np.sum(p / E_SCALE, axis=1, dtype=F32)
        if want_peaks:
            peak_off[d] = np.argmax(p, axis=1)
            peak_val[d] = p[np.arange(M), peak_off[d]]
    if sum_all_masks:
        Es = np.zeros_like(E)
        acc = E[:, 0].copy()
        for m in range(1, M):
            acc = (acc + E[:, m]).astype(F32)
        Es[:, 0] = acc
        E = Es
    if want_peaks:
        return E, peak_val, peak_off
    return E


def search_energy_parseval(X, masks, shifts, sum_all_masks):
    """Parseval form of the same metric: no inverse FFT.
    E[d,m] = (N / 2^18) * sum_k abs2(X[(k+s_d)%N]) * abs2(Mk[m,k])."""
    D, (M, N) = len(shifts), masks.shape
    PX = abs2(X).astype(np.float64)
    PM = abs2(masks).astype(np.float64)
    E = np.zeros((D, M), dtype=F32)
    for d in range(D):
        E[d] = (PM @ np.roll(PX, -int(shifts[d])) * (N / 262144.0)).astype(F32)
    if sum_all_masks:
        Es = np.zeros_like(E)
        Es[:, 0] = E.sum(axis=1, dtype=np.float64).astype(F32)
        E = Es
    return E


# --------------------------------------------------------------------------------------
# a9: findDopplerEst (kern:502-597)
# --------------------------------------------------------------------------------------
def top2_scan(col, start, count):
    """Running top-2 of ``col[start:start+count]`` exactly as kern:527-544.
    Returns (maxVal[2], maxIdx[2], cur) with float32 values."""
    maxVal = [F32(0), F32(0)]
    maxIdx = [0, 0]
    cur = 0
    for i in range(start, start + count):
        v = col[i]
        if v > maxVal[cur]:
            maxVal[cur] = v
            maxIdx[cur] = i
            cur = 1 if maxVal[0] >= maxVal[1] else 0
    return maxVal, maxIdx, cur


def find_doppler_est(E, num_dopplers, element_offset, sum_all_masks):
    """Returns float32 [best index, 10*log10(metric)] (kern:546-596)."""
    E = np.asarray(E, dtype=F32)
    M = E.shape[1]
    cols = [0] if sum_all_masks else range(M)
    idxL, valL = [], []
    with np.errstate(all="ignore"):
        for m in cols:
            v, i, cur = top2_scan(E[:, m], element_offset, num_dopplers)
            tmp = _fma32(F32(i[0]), v[0], F32(i[1]) * v[1])            # SASS: FMUL + FFMA
            idxL.append(F32(tmp / F32(v[0] + v[1])))
            if element_offset > 0:
                valL.append(F32(v[(cur + 1) % 2] / E[0, m]))             # kern:550-554
            else:
                valL.append(F32(tmp / F32(i[0] + i[1])))
        if sum_all_masks:
            best, metric = idxL[0], valL[0]
        else:
            # butterfly shuffle sum over the M active lanes (kern:488-499, :576-580)
            best = F32(_lane_sum(idxL) / F32(M))
            metric = F32(_lane_sum(valL) / F32(M))
        return np.array([best, F32(10) * np.log10(metric, dtype=F32)], dtype=F32)


def _lane_sum(vals):
    """xor-butterfly sum as activeWarpReduceSum does for a power-of-two lane count."""
    v = [F32(x) for x in vals]
    n = len(v)
    if n & (n - 1):
        return F32(np.sum(np.array(v, dtype=F32)))
    step = n >> 1
    while step:
        v = [F32(v[i] + v[i ^ step]) for i in range(n)]
        step >>= 1
    return v[0]


# --------------------------------------------------------------------------------------
# a5: host interpolation (dem_base:604-632)
# --------------------------------------------------------------------------------------
def interpolate_doppler(best, shifts, doppHzLUT):
    """lowIdx, highIdx, interpolated Hz (float64), interpolated shift (int32).
    Raises ValueError for NaN like ``int(nan)`` does in the reference (dem_base:610,625)."""
    b = np.float64(best)
    if math.isnan(b):
        raise ValueError("cannot convert float NaN to integer")
    lowIdx = int(b)
    highIdx = int(np.ceil(b))
    frac = b % 1
    hz = doppHzLUT[lowIdx] + (doppHzLUT[highIdx] - doppHzLUT[lowIdx]) * frac
    s_lo, s_hi = int(shifts[lowIdx]), int(shifts[highIdx])
    shift = np.int32(np.round(s_lo + (s_hi - s_lo) * frac))
    return lowIdx, highIdx, hz, shift


# --------------------------------------------------------------------------------------
# a10: computeSNR (dem_base:635-667)
# --------------------------------------------------------------------------------------
def compute_snr(X, shifts, lowIdx, highIdx, windowWidth, Nfft):
    lo = int(shifts[lowIdx])
    hi = int(shifts[highIdx])
    nlo = (lo + Nfft // 2) % Nfft
    nhi = (hi + Nfft // 2) % Nfft

    def window_mean(a, b):
        if a > b:
            seg = np.concatenate((np.abs(X[a - windowWidth:]), np.abs(X[:b + windowWidth])))
        else:
            seg = np.abs(X[a - windowWidth:b + windowWidth])
        return np.mean(seg)

    with np.errstate(all="ignore"):
        sig = window_mean(lo, hi)
        noise = window_mean(nlo, nhi)
        ratio = F32(sig) / F32(noise)                # float32 / float32 (dem_base:663)
        return np.float64(20) * np.log10(np.float64(ratio) - 1)


# --------------------------------------------------------------------------------------
# a19: __thresholdInput (dem_base:670-707) -- STX backend only
# --------------------------------------------------------------------------------------
def threshold_input(samples, peakThresholdScale, Nfft, peakMinGap=100):
    """In-place two-pass clipping. Returns (clippedPeakIPure, clippedPeakI)."""
    a = np.abs(samples)
    thresh = peakThresholdScale * np.mean(a)
    i = np.where(a > thresh)[0]
    samples[i] = thresh * (samples[i] / a[i])
    a[i] = np.abs(samples[i])
    thresh = peakThresholdScale * np.mean(a)
    i = np.where(a > thresh)[0]
    pure = i
    samples[i] = thresh * (samples[i] / a[i])
    if len(pure) > 0:
        diffPeaks = np.diff(pure)
        gapsAll = np.where(diffPeaks > 1)[0]
        gaps = np.where(diffPeaks[gapsAll] < peakMinGap)[0]
        gapsLen = diffPeaks[gapsAll[gaps]]
        gapsIdx = gapsAll[gaps]
        pp = np.zeros(Nfft, dtype=np.int8)
        pp[pure] = 1
        for k in range(len(gapsLen)):
            pp[pure[gapsIdx[k]]:pure[gapsIdx[k]] + gapsLen[k]] = 1
        filled = np.where(pp == 1)[0]
    else:
        filled = pure.copy()
    return pure, filled


# --------------------------------------------------------------------------------------
# a13/a14: timing recovery (kern:191-205, dem_base:717-752, kern:236-320)
# --------------------------------------------------------------------------------------
def sum_masks_abs2(y, code_search_mask_offset=0):
    """p[n] = sum over masks of abs2(y[m, n]), accumulated in mask order (kern:191-205)."""
    M = y.shape[0]
    p = np.zeros(y.shape[1], dtype=F32)
    for m in range(code_search_mask_offset, M - code_search_mask_offset):
        p = (p + abs2(y[m])).astype(F32)
    return p


def code_rate_band(Nfft, spsym):
    """(offsetHigh, offsetLow) = (int(N/(1.1 sps)), int(N/(0.9 sps)))  (dem_base:508-512)."""
    return int(Nfft / (1.1 * spsym)), int(Nfft / (0.9 * spsym))


def find_code_rate_and_phase(Pf, offset, length, abs2=abs2):
    """[index of max, atan2 at max, max abs2] over Pf[offset:offset+length] (kern:236-320).
    Ties: lowest index (the reference is lane-order dependent on exact ties)."""
    band = abs2(Pf[offset:offset + length])
    idx = int(np.argmax(band)) + offset
    v = Pf[idx]
    return np.array([F32(idx), np.arctan2(F32(v.imag), F32(v.real)), abs2(v)], dtype=F32)


def code_rate_host(result, Nfft):
    """spSym and codeOffset from the kernel result, float64 host math (dem_base:733-752)."""
    with np.errstate(all="ignore"):
        spSym = np.float64(Nfft) / np.float64(result[0])
        codeOffset = -np.float64(result[1]) / np.pi * spSym / 2
        if codeOffset < 0:
            codeOffset += spSym - 1
    return spSym, codeOffset


# --------------------------------------------------------------------------------------
# a15: findCentres (kern:78-146, dem_base:991-1009)
# --------------------------------------------------------------------------------------
def find_centres(ymag2, spSym, codeOffset, Nfft, windowWidth, spsymMin):
    """Per-symbol decision. ``ymag2`` = abs2(y) float32 [M, N].
    Returns (symbols int32[S], centres int32[S], magnitudes float32[S])."""
    if spSym < spsymMin:                       # dem_base:994-995
        spSym = spsymMin
    S = int(Nfft / spSym)                      # dem_base:999
    sp32 = F32(spSym)
    off32 = F32(codeOffset)
    M = ymag2.shape[0]
    W = windowWidth
    left = W // 2
    x = np.arange(S, dtype=np.int64)
    t = (x.astype(np.float64) * np.float64(sp32) - np.float64(left)).astype(F32)   # FFMA
    a0 = np.trunc((t + off32).astype(F32)).astype(np.int64)
    oc = np.full(S, int(np.trunc(off32)), dtype=np.int64)
    neg = a0 < 0
    oc[neg] -= a0[neg]
    a = np.where(neg, 0, a0)
    e = np.minimum(a0 + W, Nfft) - a                                             # kern:99-102
    k = np.arange(W)[None, :]
    idx = np.minimum(a[:, None] + k, Nfft - 1)
    win = ymag2[:, idx]                                                          # [M, S, W]
    win = np.where(k[None, :, :] < e[None, :, None], win, F32(-1))
    flat = np.transpose(win, (1, 0, 2)).reshape(S, M * W)                        # mask-major
    j = np.argmax(flat, axis=1)
    best = flat[np.arange(S), j]
    hit = best > 0
    sym = np.where(hit, j // W, -1).astype(np.int32)
    kc = np.where(hit, j % W, -1)
    mag = np.where(hit, best, F32(0)).astype(F32)
    c = (t + kc.astype(F32)).astype(F32)
    c = (c + oc.astype(F32)).astype(F32)
    centres = np.trunc(c).astype(np.int32)
    return sym, centres, mag


def trust_from_magnitudes(mag, S):
    """The reference reads the float magnitudes back through an int8-sized buffer
    (dem_base:472,1005-1007): trust[i] = i-th raw byte of the float32 array."""
    return np.ascontiguousarray(mag, dtype=F32).view(np.int8)[:S].copy()


# --------------------------------------------------------------------------------------
# a16: extractBits / extractBitsNRZs (dem_base:1012-1051)
# --------------------------------------------------------------------------------------
def extract_bits(symbols, bitLUT, symbolLUT):
    if bitLUT is None:
        if len(np.shape(symbolLUT)) == 3:
            return extract_bits_nrzs(symbols, symbolLUT)
        raise NotImplementedError("extractBitsOld is undefined in the reference (dem_base:1017)")
    return np.asarray(bitLUT)[symbols], []


def extract_bits_nrzs(symbols, symbolLUT):
    nxt = symbols[1:, None]
    res1 = np.any(nxt == symbolLUT[symbols[:-1], 0, :], axis=1)
    res0 = np.any(nxt == symbolLUT[symbols[:-1], 1, :], axis=1)
    symError = np.where((res1 + res0) == 0)[0].tolist()
    res1[symError] = int(SYMBOL_MISMATCHVAL)
    return res1, symError


# --------------------------------------------------------------------------------------
# a17: checkSymbolOverlap (dem_base:863-988)
# --------------------------------------------------------------------------------------
class OverlapState:
    """Cross-chunk state: bits after the window end and the last bits inside the window."""

    def __init__(self):
        self.poswinP = []
        self.posSymEnd = None


def _eq(a, b):
    """Elementwise == that degrades to a scalar False on a length mismatch, as NumPy < 1.25
    did (newer NumPy raises, which the reference's blanket ``except`` turns into "no change")."""
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return np.bool_(False)
    return a == b


def check_symbol_overlap(state, noError, centres, dataBits, trustSymbol, Nfft, sigOverlapWin,
                         overlapOffset, error_threshold, match_threshold):
    """Returns (centresWin, dataBitsWin, trustSymbolWin) and updates ``state``."""
    startOverlap = np.where(centres >= sigOverlapWin)[0][0]
    endOverlap = np.where(centres > (Nfft - sigOverlapWin))[0][0]
    win = dataBits[startOverlap:endOverlap]
    pre = dataBits[:startOverlap]
    oo = overlapOffset
    if noError > error_threshold:
        pass
    elif len(state.poswinP) > 0:
        P, Eend = state.poswinP, state.posSymEnd
        if np.all(_eq(P[:oo], win[:oo])) or np.all(_eq(Eend[-oo:], pre[-oo:])):
            pass
        else:
            symPre = np.sum(_eq(P[:oo], win[:oo]))
            symPos = np.sum(_eq(Eend[-oo:], pre[-oo:]))
            symEarlyPre = np.sum(_eq(P[:oo], win[1:oo + 1]))
            symEarlyPos = np.sum(_eq(Eend[-oo - 1:-1], pre[-oo:]))
            symLatePre = np.sum(_eq(P[1:oo + 1], win[0:oo]))
            symLatePos = np.sum(_eq(Eend[-oo:], pre[-oo - 1:-1]))
            maxPre = max(symPre, symEarlyPre, symLatePre)
            maxPos = max(symPos, symEarlyPos, symLatePos)
            if match_threshold < symEarlyPre and symEarlyPre == maxPre:
                if match_threshold < symEarlyPos and symEarlyPos == maxPos:
                    startOverlap += 1
            elif match_threshold < symLatePre and symLatePre == maxPre:
                if match_threshold < symLatePos and symLatePos == maxPos:
                    startOverlap -= 1
    dataBitsWin = dataBits[startOverlap:endOverlap]
    state.poswinP = dataBits[endOverlap:]
    state.posSymEnd = dataBitsWin[-oo - 1:]
    return centres[startOverlap:endOverlap], dataBitsWin, trustSymbol[startOverlap:endOverlap]


# --------------------------------------------------------------------------------------
# a11/a18: clipped-peak tagging and output casts (dem_base:817-859)
# --------------------------------------------------------------------------------------
def tag_clipped_peaks(trustWin, centresWin, clippedPeakIPure, spSym, Nfft):
    pp = np.zeros(Nfft, dtype=bool)
    spSymc = int(np.ceil(spSym))
    for cp in clippedPeakIPure:
        pp[cp - 2 * spSymc:cp + 2 * spSymc + 1] = 1
    trustWin = trustWin.copy()
    trustWin[pp[centresWin]] = -2
    return trustWin


# --------------------------------------------------------------------------------------
# Doppler-rate dimension: complexHeterodyne, prepared at dem_base:388 and never called
# --------------------------------------------------------------------------------------
def heterodyne(x, a, b=0.0, c=0.0):
    """out[n] = x[n] * exp(j theta), theta = fmod(((a n) + b) n, 2 pi) + c, everything in float32 (kern:755-778; the
    (a n) + b is contracted to one fma by nvcc)."""
    n = np.arange(len(x), dtype=np.float64)
    ab = (np.float64(F32(a)) * n + np.float64(F32(b))).astype(F32)          # fma(a, n, b): exact in float64, rounded once
    theta = (ab.astype(np.float64) * n).astype(F32)
    theta = (np.fmod(theta, F32(2.0) * F32(np.pi)) + F32(c)).astype(F32)
    w = (np.cos(theta).astype(F32) + 1j * np.sin(theta).astype(F32)).astype(np.complex64)
    return (np.asarray(x, dtype=np.complex64) * w).astype(np.complex64)


def rate_to_a(rate_hz_per_s, fs):
    """De-chirp coefficient for a linear Doppler rate r [Hz/s]: the signal carries exp(+j pi r t^2) (signals.doppler_rate)."""
    return F32(-np.pi * float(rate_hz_per_s) / float(fs) ** 2)


def search_rates(x, masks, shifts, sum_all_masks, rates_hz_per_s, fs, element_offset=0, workers=1):
    """One ordinary Doppler search per rate hypothesis on the de-chirped chunk; metric = the largest mask-summed energy.
    Returns (index of the best rate -- first one on ties --, metrics[rate], E[rate])."""
    Es, metrics = [], []
    for r in rates_hz_per_s:
        X = forward_fft(heterodyne(x, rate_to_a(r, fs)))
        E = search_energy(X, masks, shifts, sum_all_masks, workers=workers)
        Es.append(E)
        metrics.append(float(np.max(np.sum(E[element_offset:], axis=1, dtype=np.float64))))
    return int(np.argmax(metrics)), np.array(metrics), Es


# --------------------------------------------------------------------------------------
# The whole per-chunk path behind the reference's class contract
# --------------------------------------------------------------------------------------
class OracleDemodulator:
    """NumPy twin of ``demodulator.UHF.Demodulator`` / ``demodulator.STX.Demodulator``."""

    def __init__(self, conf, protocol, radioName, backend="UHF", fft_workers=1):
        self.backend = backend
        self.fft_workers = fft_workers
        self.protocol = protocol
        self.confRadio = cr = conf["Radios"]["Rx"][radioName]
        self.confGPU = cg = conf["GPU"][cr["CUDA_settings"]]
        self.Nfft = 2 ** cg["blockSize"]
        self.sigOverlap = 2 ** cg["overlap"]
        self.sigOverlapWin = int(self.sigOverlap / 2)
        self.clippedPeakSpan = cg["clippedPeakSpan"]
        self.peakThresholdScale = cg["peakThresholdScale"]
        self.overlapOffset = cg.get("symbol_check_overlap_offset", SYMBOL_CHECK_OVERLAP_OFFSET)
        self.symbol_check_error_threshold = cg.get("symbol_check_error_threshold",
                                                   SYMBOL_CHECK_ERROR_THRESHOLD)
        self.symbol_check_match_threshold = self.overlapOffset - cg.get(
            "symbol_check_match_num_errors_allowed", SYMBOL_CHECK_MATCH_NUM_ERRORS_ALLOWED)
        self.spsym = cr["samplesPerSym"]
        self.spsymMin = int(self.spsym / 2)
        self.baudRate = cr["baud"]
        self.sampleRate = self.baudRate * self.spsym
        self.windowWidth = cg["bitWindowWidth"]
        self.CODE_SEARCH_MASK_OFFSET = 0
        self.SUM_ALL_MASKS_PYTHON = bool(getattr(protocol, "SUM_ALL_MASKS_PYTHON", False))
        g = doppler_grid(conf, radioName, self.Nfft)
        self.grid = g
        self.num_dopplers = g["num_dopplers"]
        self.doppIdxArrayOffset = g["element_offset"]
        self.doppCyperSymNorm = g["shifts"]
        self.doppHzLUT = g["doppHzLUT"]
        self.doppOffsetIdx = g["doppOffsetIdx"]
        self.centreFreqOffset = g["centreFreqOffset"]
        self.num_masks, masks = protocol.get_filter(self.Nfft, self.spsym, cg["xcorrMaskSize"])
        if masks.shape != (self.num_masks, self.Nfft):                       # dem_base:252-255
            raise ValueError("Masks provided by protocol {} expected to be of dimensions {}, got dimensions {}".format(
                protocol.name, (self.num_masks, self.Nfft), masks.shape))
        if not isinstance(masks[0, 0], np.complex64):                        # dem_base:256-257
            raise TypeError("Datatype of masks {}, expected {}".format(type(masks[0, 0]), np.complex64))
        self.masks = masks
        self.bitLUT, self.symbolLUT = protocol.get_symbolLUT2(cg["xcorrMaskSize"])
        self.iHigh, self.iLow = code_rate_band(self.Nfft, self.spsym)
        self.raw = np.zeros(self.Nfft, dtype=np.complex64)
        self.clippedPeakIPure = []
        self.state = OverlapState()
        self.dopplerIdxlast = 0
        self.last = {}

    def get_signalBufferHostPointer(self):
        return self.raw

    def uploadAndFindCarrier(self, samples):
        if self.backend == "STX":                                            # STX.py:8-25
            self.clippedPeakIPure, _ = threshold_input(samples, self.peakThresholdScale, self.Nfft)
            self.X = forward_fft(samples)
            return 0, 0, self.clippedPeakIPure, 0
        self.X = forward_fft(samples)                                        # UHF.py:15
        E = search_energy(self.X, self.masks, self.doppCyperSymNorm, self.SUM_ALL_MASKS_PYTHON,
                          workers=self.fft_workers)
        res = find_doppler_est(E, self.num_dopplers, self.doppIdxArrayOffset, self.SUM_ALL_MASKS_PYTHON)
        self.last.update(E=E, res=res)
        try:
            lo, hi, hz, shift = interpolate_doppler(res[0], self.doppCyperSymNorm, self.doppHzLUT)
            self.dopplerIdxlast = shift
            SNR = compute_snr(self.X, self.doppCyperSymNorm, lo, hi, 5, self.Nfft)
            freqOffset = hz - self.centreFreqOffset
            sdev_Hz = res[1] / self.Nfft * self.sampleRate
        except ValueError:
            self.dopplerIdxlast = 0
            freqOffset, sdev_Hz, SNR = 0.0, 0.0, 0.0
        return freqOffset, sdev_Hz, self.clippedPeakIPure, SNR

    def demodulate(self):
        if self.backend == "STX":
            self.dopplerIdxlast = self.doppOffsetIdx                         # dem_base:758-761
        y = surface_rows(self.X, self.masks, self.dopplerIdxlast)
        p = sum_masks_abs2(y, self.CODE_SEARCH_MASK_OFFSET)
        Pf = scipy.fft.rfft(p)
        r = find_code_rate_and_phase(Pf, self.iHigh, self.iLow - self.iHigh)
        spSym, codeOffset = code_rate_host(r, self.Nfft)
        ymag2 = abs2(y)
        sym, centres, mag = find_centres(ymag2, spSym, codeOffset, self.Nfft, self.windowWidth, self.spsymMin)
        trust = trust_from_magnitudes(mag, len(sym))
        self.last.update(shift=int(self.dopplerIdxlast), timing=r, spSym=spSym, codeOffset=codeOffset,
                         sym=sym, centres=centres, mag=mag, y=y)
        dataBits, symErr = extract_bits(sym, self.bitLUT, self.symbolLUT)
        cW, bW, tW = check_symbol_overlap(self.state, len(symErr), centres, dataBits, trust, self.Nfft,
                                          self.sigOverlapWin, self.overlapOffset,
                                          self.symbol_check_error_threshold, self.symbol_check_match_threshold)
        tW = tag_clipped_peaks(tW, cW, self.clippedPeakIPure, spSym, self.Nfft)
        return bW.astype(np.uint8), cW.astype(np.uint8), tW.astype(np.uint8), spSym


def run_stream(demod, samples, overlap=None):
    """Feed ``samples`` through ``demod`` chunk by chunk exactly like the reference's process
    loop (demodulator_process.py:284-338): raw[ovl:] = new block; search; demodulate;
    raw[:ovl] = raw[-ovl:].  Works for the oracle and for the CUDA-backed class alike.
    Returns a list of per-chunk dicts."""
    N = demod.Nfft
    ovl = demod.sigOverlap if overlap is None else overlap
    step = N - ovl
    raw = demod.get_signalBufferHostPointer()
    raw[:] = 0
    out = []
    for c in range(len(samples) // step):
        raw[ovl:] = samples[c * step:(c + 1) * step]
        doppler, doppler_std, clipIdx, SNR = demod.uploadAndFindCarrier(raw)
        bits, centres, trust, spSym = demod.demodulate()
        out.append({"doppler": doppler, "doppler_std": doppler_std, "SNR": SNR, "data": bits,
                    "centres": centres, "trust": trust, "spSymEst": spSym})
        raw[:ovl] = raw[-ovl:]
    return out
