"""ctypes binding of libpycusdr_b200.so (include/pycusdr_b200.h).

The library is built in-tree by ``pycusdr_b200/csrc/Makefile`` (or ``__graft_entry__.build()``).
There is no CPU fallback: if the shared object is missing or no CUDA device is usable the
caller gets an exception, never a silently slower path.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PYCUSDR_B200_LIB") or os.path.join(_HERE, "libpycusdr_b200.so")     # (env: kernel experiments)
ABI_VERSION = 1

PATH_AUTO, PATH_OVERLAP_SAVE, PATH_FULL, PATH_PARSEVAL = 0, 1, 2, 3


class NativeError(RuntimeError):
    """A pcs_* call failed; the message is pcs_last_error()."""

    def __init__(self, code, msg):
        super().__init__(f"pycusdr_b200 native error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "device", "nfft", "num_dopplers", "element_offset", "num_masks", "window_width",
        "sum_all_masks", "code_search_mask_offset", "samples_per_sym", "path", "log2_block", "snr_window")] + [
        ("reserved", C.c_int32 * 3)]


class Result(C.Structure):
    _fields_ = [
        ("best_idx", C.c_float), ("metric_db", C.c_float), ("low_idx", C.c_int32), ("high_idx", C.c_int32),
        ("shift", C.c_int32), ("status", C.c_int32), ("timing", C.c_float * 3), ("n_sym", C.c_int32),
        ("sp_sym", C.c_double), ("code_offset", C.c_double), ("peak_val", C.c_float), ("peak_bin", C.c_int32),
        ("peak_mask", C.c_int32), ("peak_offset", C.c_int32), ("sig_start", C.c_int32), ("sig_len", C.c_int32),
        ("noise_start", C.c_int32), ("noise_len", C.c_int32), ("demod_shift", C.c_int32), ("xchg_timeout", C.c_int32)]


class StitchConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("nfft", "overlap", "overlap_offset", "error_threshold", "match_threshold",
                                         "num_symbols", "lut_k", "reserved")]


class PlanInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "path", "log2_block", "valid_per_block", "num_blocks", "support_pos", "support_neg", "groups_per_cta",
        "search_ctas", "search_smem_bytes", "sm_count")] + [("device_bytes", C.c_int64)]


# every symbol include/pycusdr_b200.h declares: (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "pcs_create": (C.c_int, [C.POINTER(Config), _P, _P, C.POINTER(_P)]),
    "pcs_destroy": (C.c_int, [_P]),
    "pcs_host_buffer": (_P, [_P]),
    "pcs_upload": (C.c_int, [_P]),
    "pcs_upload_thresholded": (C.c_int, [_P, C.c_float, _P, C.c_int32, C.POINTER(C.c_int32), _P]),
    "pcs_upload_device": (C.c_int, [_P, _P]),
    "pcs_heterodyne": (C.c_int, [_P, C.c_float, C.c_float, C.c_float]),
    "pcs_get_chunk": (C.c_int, [_P, _P]),
    "pcs_search": (C.c_int, [_P, C.POINTER(Result), _P]),
    "pcs_demod": (C.c_int, [_P, C.c_int32, C.POINTER(Result), _P, _P, _P]),
    "pcs_process": (C.c_int, [_P, C.POINTER(Result), _P, _P, _P, _P]),
    "pcs_enqueue_device": (C.c_int, [_P, _P]),
    "pcs_fetch": (C.c_int, [_P, C.POINTER(Result), _P, _P, _P, _P]),
    "pcs_max_symbols": (C.c_int32, [_P]),
    "pcs_snr_windows": (C.c_int, [_P, _P, _P]),
    "pcs_get_spectrum": (C.c_int, [_P, _P]),
    "pcs_get_peaks": (C.c_int, [_P, _P, _P]),
    "pcs_get_demod_surface": (C.c_int, [_P, C.c_int32, _P]),
    "pcs_get_demod_magnitudes": (C.c_int, [_P, _P, _P]),
    "pcs_get_plan": (C.c_int, [_P, C.POINTER(PlanInfo)]),
    "pcs_factorise_bank": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int32, C.c_int32,
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P, _P, _P]),
    "pcs_bank_code_order": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, C.c_int32, C.POINTER(C.c_int32)]),
    "pcs_get_bank_factor": (C.c_int, [_P, _P]),
    "pcs_launch_count": (C.c_int64, [_P]),
    "pcs_stream": (C.c_uint64, [_P]),
    "pcs_set_bin_range": (C.c_int, [_P, C.c_int32, C.c_int32]),
    "pcs_shard_buffers": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "pcs_enqueue_search_local": (C.c_int, [_P]),
    "pcs_enqueue_estimate_and_demod": (C.c_int, [_P, C.c_int32]),
    "pcs_shard_init": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "pcs_shard_attach": (C.c_int, [_P, _P]),
    "pcs_shard_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_int32)]),
    "pcs_shard_host_slot": (C.c_int, [_P, C.c_int64, C.POINTER(_P)]),
    "pcs_shard_submit": (C.c_int, [_P, C.c_int64, C.c_int32, _P]),
    "pcs_shard_fetch": (C.c_int, [_P, C.c_int64, C.POINTER(Result), _P, _P, _P, _P, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                  C.POINTER(C.c_int32)]),
    "pcs_shard_sync": (C.c_int, [_P]),
    "pcs_shard_trace": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int32), _P]),
    "pcs_shard_streams": (C.c_int, [_P, _P]),
    "pcs_stitch_create": (C.c_int, [C.POINTER(StitchConfig), _P, _P, C.POINTER(_P)]),
    "pcs_stitch_chunk": (C.c_int, [_P, _P, _P, _P, C.c_int32, _P, C.c_int32, C.c_double, _P, _P, _P, C.POINTER(C.c_int32)]),
    "pcs_stitch_reset": (C.c_int, [_P]),
    "pcs_stitch_get_state": (C.c_int, [_P, _P, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "pcs_stitch_set_state": (C.c_int, [_P, _P, C.c_int32, C.c_int32]),
    "pcs_stitch_destroy": (C.c_int, [_P]),
    "pcs_ingest_create": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "pcs_ingest_push": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int32)]),
    "pcs_ingest_pop": (C.c_int, [_P, C.c_int32, C.POINTER(Result), _P, _P, _P, _P, _P, _P, C.POINTER(C.c_int32)]),
    "pcs_ingest_pending": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "pcs_ingest_destroy": (C.c_int, [_P]),
    "pcs_snr_means": (C.c_int, [_P, _P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "pcs_mean_abs_c64": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_float)]),
    "pcs_fill_gaps": (C.c_int, [_P, C.c_int32, C.c_int32, _P, C.c_int32, C.POINTER(C.c_int32)]),
    "pcs_chunk_to_bits": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.POINTER(Result), _P, _P, _P, _P, C.POINTER(C.c_float),
                                    C.POINTER(C.c_float), C.POINTER(C.c_int32), _P, _P, _P, C.POINTER(C.c_int32)]),
    "pcs_sync_search": (C.c_int, [_P, C.c_int64, _P, C.c_int32, C.c_int32, _P, _P, C.c_int32, C.POINTER(C.c_int32)]),
    "pcs_bit_xcorr": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int64, _P]),
    "pcs_topk_i32": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P]),
    "pcs_host_register": (C.c_int, [_P, C.c_uint64]),
    "pcs_host_unregister": (C.c_int, [_P]),
    "pcs_set_host_source": (C.c_int, [_P, _P]),
    "pcs_set_stream": (C.c_int, [_P, C.c_uint64]),
    "pcs_set_profiling": (C.c_int, [_P, C.c_int]),
    "pcs_get_profile": (C.c_int, [_P, _P, _P]),
    "pcs_measure_fp32_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "pcs_last_error": (C.c_char_p, []),
    "pcs_abi_version": (C.c_int, []),
}

_lib = None
_live = None      # weak set of Engine / Stitcher objects, closed at interpreter exit while CUDA is still up


def load():
    """Load the shared library (once) and type its entry points. Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `make -C pycusdr_b200/csrc` "
                "(or __graft_entry__.build()); pycusdr_b200 has no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.pcs_abi_version() != ABI_VERSION:
            raise ImportError(f"ABI mismatch: library {lib.pcs_abi_version()}, binding {ABI_VERSION}")
        _lib = lib
        import atexit
        import weakref
        global _live
        _live = weakref.WeakSet()

        def _close_all():
            for obj in list(_live):
                try:
                    obj.close()
                except Exception:
                    pass
        atexit.register(_close_all)
    return _lib


def measure_fp32_peak(device=0):
    """Measured fp32 FMA throughput of ``device`` in TFLOP/s."""
    lib = load()
    out = C.c_double(0)
    rc = lib.pcs_measure_fp32_peak(int(device), C.byref(out))
    if rc != 0:
        raise NativeError(rc, lib.pcs_last_error().decode())
    return out.value


def mean_abs_c64(z):
    """float32 mean of |z| for a complex64 window (``pcs_mean_abs_c64``; host only)."""
    z = np.ascontiguousarray(z, dtype=np.complex64)
    out = C.c_float(0)
    rc = load().pcs_mean_abs_c64(z.__array_interface__["data"][0], len(z), C.byref(out))
    if rc != 0:
        raise NativeError(rc, load().pcs_last_error().decode())
    return np.float32(out.value)


FB_MAX_SEG, FB_MAX_BASIS = 4, 4


def factorise_bank(masks, support_pos, support_neg, shifts, log2_block):
    """Segment factorisation of a filter bank (``pcs_factorise_bank``; host only, no GPU).

    ``masks``: complex64[M, nfft] spectra as given to ``pcs_create``.  Returns ``None`` when the bank does not factorise
    (or it would not pay), else a dict with ``S`` (segment length), ``J`` (segments per filter), ``R`` (basis filters),
    ``sel`` int32[M, J], ``coef`` complex64[D, M, J] and ``basis_spec`` complex64[D, R, 2**log2_block].
    """
    masks = np.ascontiguousarray(masks, dtype=np.complex64)
    shifts = np.ascontiguousarray(shifts, dtype=np.int32)
    M, N = masks.shape
    D, B = len(shifts), 1 << int(log2_block)
    sel = np.zeros(M * FB_MAX_SEG, dtype=np.int32)
    coef = np.zeros(D * M * FB_MAX_SEG, dtype=np.complex64)
    spec = np.zeros(D * FB_MAX_BASIS * B, dtype=np.complex64)
    S, J, R = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    rc = load().pcs_factorise_bank(_ptr(masks), N, M, int(support_pos), int(support_neg), _ptr(shifts), D, int(log2_block),
                                   C.byref(S), C.byref(J), C.byref(R), _ptr(sel), _ptr(coef), _ptr(spec))
    if rc != 0:
        raise NativeError(rc, load().pcs_last_error().decode())
    if R.value == 0:
        return None
    S, J, R = S.value, J.value, R.value
    return {"S": S, "J": J, "R": R, "sel": sel[:M * J].reshape(M, J).copy(),
            "coef": coef[:D * M * J].reshape(D, M, J).copy(), "basis_spec": spec[:D * R * B].reshape(D, R, B).copy()}


def bank_code_order(fact, allow_shared_sums=True):
    """Form of the factorised search a bank can take (``pcs_bank_code_order``; host only): returns ``(form, sel, coef)`` --
    for form >= 2 ``coef`` is in code order and ``sel`` is the table code -> mask."""
    sel = np.ascontiguousarray(fact["sel"], dtype=np.int32).copy()
    coef = np.ascontiguousarray(fact["coef"], dtype=np.complex64).copy()
    D, M, J = coef.shape
    form = C.c_int32(0)
    rc = load().pcs_bank_code_order(M, J, int(fact["R"]), D, _ptr(sel), _ptr(coef), int(bool(allow_shared_sums)), C.byref(form))
    if rc != 0:
        raise NativeError(rc, load().pcs_last_error().decode())
    return form.value, (sel.reshape(-1)[:M].copy() if form.value >= 2 else sel), coef


def fill_gaps(idx, min_gap, nfft):
    """``clippedPeakI`` from ``clippedPeakIPure`` (dem_base:686-705, ``pcs_fill_gaps``; host only)."""
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    out = np.empty(int(nfft), dtype=np.int64)
    n = C.c_int32(0)
    rc = load().pcs_fill_gaps(_ptr(idx), len(idx), int(min_gap), _ptr(out), len(out), C.byref(n))
    if rc != 0:
        raise NativeError(rc, load().pcs_last_error().decode())
    return out[:min(n.value, len(out))].copy()


class HostRegistration:
    """A range of the caller's own host memory page-locked for direct H2D copies (``pcs_host_register``): typically the
    sample ring a receiver writes into, of which consecutive chunks are overlapping windows (sigFIFO.py:147-181).  Keeps the
    array alive; ``close()`` (or garbage collection) releases the page lock."""

    def __init__(self, arr):
        if not (isinstance(arr, np.ndarray) and arr.flags.c_contiguous and arr.nbytes > 0):
            raise ValueError("host_register needs a non-empty C-contiguous NumPy array")
        self.lib = load()
        self.arr = arr
        self.start = arr.__array_interface__["data"][0]
        self.end = self.start + arr.nbytes
        rc = self.lib.pcs_host_register(self.start, arr.nbytes)
        if rc != 0:
            raise NativeError(rc, self.lib.pcs_last_error().decode())
        _registered.append(self)
        _live.add(self)

    def contains(self, ptr, nbytes):
        return self.start <= ptr and ptr + nbytes <= self.end

    def close(self):
        if self.arr is not None:
            if self in _registered:
                _registered.remove(self)
            self.lib.pcs_host_unregister(self.start)
            self.arr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_registered = []


def host_register(arr):
    """Page-lock ``arr`` (in place, no copy) so that chunks that are views of it go to the GPU without a staging copy."""
    return HostRegistration(arr)


def registered_ptr(a, nbytes):
    """Address of ``a``'s buffer when ``a`` is a C-contiguous array of exactly ``nbytes`` bytes inside a registered range,
    else None."""
    if not isinstance(a, np.ndarray) or not _registered or a.nbytes != nbytes or not a.flags.c_contiguous:
        return None
    ptr = a.__array_interface__["data"][0]
    for r in _registered:
        if r.contains(ptr, nbytes):
            return ptr
    return None


def _ptr(a):
    """NULL or the address of a NumPy array's buffer (``a.ctypes.data_as`` costs ~3 us a call; this does not)."""
    return None if a is None else a.__array_interface__["data"][0]


class Stitcher:
    """Host-side bit post-processing of a chunk (``pcs_stitch_*``): bit extraction, chunk stitching, clipped-peak
    tagging and the uint8 casts in one native call; keeps the cross-chunk state."""

    def __init__(self, *, nfft, overlap, overlap_offset, error_threshold, match_threshold, bit_lut, symbol_lut):
        self.lib = load()
        if bit_lut is not None:
            self._bit = np.ascontiguousarray(np.asarray(bit_lut), dtype=np.uint8)
            self._sym, M, K = None, len(self._bit), 0
        else:
            lut = np.asarray(symbol_lut)
            if lut.ndim != 3 or lut.shape[1] != 2:
                raise NotImplementedError("extractBitsOld is not defined by the reference either (dem_base:1017)")
            self._sym = np.ascontiguousarray(lut, dtype=np.int32)
            self._bit, M, K = None, lut.shape[0], lut.shape[2]
        cfg = StitchConfig(int(nfft), int(overlap), int(overlap_offset), int(error_threshold), int(match_threshold), M, K, 0)
        self._h = _P()
        rc = self.lib.pcs_stitch_create(C.byref(cfg), _ptr(self._bit), _ptr(self._sym), C.byref(self._h))
        if rc != 0:
            raise NativeError(rc, self.lib.pcs_last_error().decode())
        self._cap = 0

    def __call__(self, sym, centre, mag, clipped, sp_sym):
        n = len(sym)
        if n > self._cap:
            self._cap = n
            self._bits, self._cen, self._tr = (np.empty(n, dtype=np.uint8) for _ in range(3))
        sym = np.ascontiguousarray(sym, dtype=np.int32)
        centre = np.ascontiguousarray(centre, dtype=np.int32)
        mag = np.ascontiguousarray(mag, dtype=np.float32)
        clipped = np.ascontiguousarray(clipped, dtype=np.int64)
        n_out = C.c_int32(0)
        rc = self.lib.pcs_stitch_chunk(self._h, _ptr(sym), _ptr(centre), _ptr(mag), n, _ptr(clipped), len(clipped),
                                       float(sp_sym), _ptr(self._bits), _ptr(self._cen), _ptr(self._tr), C.byref(n_out))
        if rc == -4:
            raise IndexError(self.lib.pcs_last_error().decode())       # what the reference's np.where(...)[0][0] raises
        if rc != 0:
            raise NativeError(rc, self.lib.pcs_last_error().decode())
        k = n_out.value
        return self._bits[:k].copy(), self._cen[:k].copy(), self._tr[:k].copy()

    def reset(self):
        self.lib.pcs_stitch_reset(self._h)

    STATE_BYTES = 1024      # transport buffer that holds a token of every shipped geometry (state_capacity(...) <= 200)

    @staticmethod
    def state_capacity(overlap, overlap_offset, spsym_min, window_width=63):
        """Upper bound in bytes of a ``get_state`` token for this geometry (for fixed-size transports): the carry holds
        the symbols whose centre lies beyond ``nfft - overlap/2`` (poswinP, dem_base:977; symbols are at least
        ``spsym_min`` samples apart and a centre is within ``window_width`` samples of its nominal position) and the last
        ``overlap_offset + 1`` bits of the window (posSymEnd, dem_base:979), after an 8-byte header."""
        return 8 + (int(overlap) // 2 + int(window_width)) // max(int(spsym_min), 1) + 2 + int(overlap_offset) + 1

    def get_state(self):
        """The carry to the next chunk as ``bytes``: two 4-byte counts (poswinP, posSymEnd) + the bits themselves.
        The lengths are queried first, so any geometry fits."""
        a, b = C.c_int32(0), C.c_int32(0)
        self.lib.pcs_stitch_get_state(self._h, None, 0, C.byref(a), C.byref(b))      # lengths only
        n = a.value + b.value
        buf = np.empty(max(n, 1), dtype=np.uint8)
        rc = self.lib.pcs_stitch_get_state(self._h, _ptr(buf), n, C.byref(a), C.byref(b))
        if rc != 0:
            raise NativeError(rc, self.lib.pcs_last_error().decode())
        return int(a.value).to_bytes(4, "little") + int(b.value).to_bytes(4, "little") + buf[:n].tobytes()

    def set_state(self, token):
        a, b = int.from_bytes(token[:4], "little"), int.from_bytes(token[4:8], "little")
        if len(token) < 8 + a + b:
            raise ValueError(f"truncated stitcher state: {len(token)} bytes for counts {a} + {b}")
        buf = np.frombuffer(token, dtype=np.uint8, count=a + b, offset=8).copy() if a + b else np.empty(0, np.uint8)
        rc = self.lib.pcs_stitch_set_state(self._h, _ptr(buf) if a + b else None, a, b)
        if rc != 0:
            raise NativeError(rc, self.lib.pcs_last_error().decode())

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._h = None
            self.lib.pcs_stitch_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sync_search(bits, mask, threshold):
    """Frame-sync candidates of a demodulated bit stream, exactly as ``decoder.py:96-104`` computes them with
    ``np.convolve(bits, mask)``: returns ``(idxCand, score[idxCand])`` (int32 arrays)."""
    lib = load()
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    mask = np.asarray(mask)
    if mask.size and np.max(np.abs(mask)) > 127:
        raise ValueError("mask entries must fit int8 (protocol.get_mask() returns +-1)")
    mask = np.ascontiguousarray(mask, dtype=np.int8)
    cap = max(16, len(bits) // 8)
    while True:
        idx = np.empty(cap, dtype=np.int32)
        sc = np.empty(cap, dtype=np.int32)
        n = C.c_int32(0)
        rc = lib.pcs_sync_search(_ptr(bits), len(bits), _ptr(mask), len(mask), int(threshold), _ptr(idx), _ptr(sc), cap,
                                 C.byref(n))
        if rc != 0:
            raise NativeError(rc, lib.pcs_last_error().decode())
        if n.value <= cap:
            return idx[:n.value].copy(), sc[:n.value].copy()
        cap = n.value


def bit_xcorr(a, b, n=None):
    """Exact circular cross-correlation of two 0/1 streams zero padded to ``n`` (default: the reference's
    ``2**ceil(log2(len(a)))``): ``out[k] = sum_j a[(j + k) % n] * b[j]`` as int32 -- what
    ``np.abs(customXCorr(a_padded, b))`` (lib/customXCorr.py:5-17) approximates in floating point."""
    lib = load()
    a = np.ascontiguousarray(np.asarray(a) != 0, dtype=np.uint8)
    b = np.ascontiguousarray(np.asarray(b) != 0, dtype=np.uint8)
    if n is None:
        n = int(2 ** np.ceil(np.log2(max(len(a), 1))))
    n = int(n)
    out = np.empty(n, dtype=np.int32)
    rc = lib.pcs_bit_xcorr(_ptr(a), len(a), _ptr(b), len(b), n, _ptr(out))
    if rc != 0:
        raise NativeError(rc, lib.pcs_last_error().decode())
    return out


def align_bits(slave_bits, master_bits, var_multiplier, n_idx=15):
    """The soft combiner's alignment test of one slave stream against the master (softCombiner.py:697-722): returns
    ``(offset, accepted, idx, val, cond)`` -- ``idx`` / ``val`` the ``n_idx`` largest correlation values by repeated
    arg-max, ``cond = mean(val[2:]) + var_multiplier * std(val[2:])``, ``accepted = val[0] > cond``, ``offset = idx[0]``
    (where the master's bits start in the slave's)."""
    lib = load()
    n = len(slave_bits)
    corr = bit_xcorr(slave_bits, np.asarray(master_bits)[:n])
    idx = np.empty(n_idx, dtype=np.int64)
    val = np.empty(n_idx, dtype=np.int32)
    rc = lib.pcs_topk_i32(_ptr(corr), len(corr), int(n_idx), _ptr(idx), _ptr(val))
    if rc != 0:
        raise NativeError(rc, lib.pcs_last_error().decode())
    valf = val.astype(np.float64)
    cond = np.mean(valf[2:]) + var_multiplier * np.std(valf[2:])
    return int(idx[0]), bool(valf[0] > cond), idx, valf, cond


class Ingest:
    """Native sample ingest over several engines (``pcs_ingest_*``): push samples in any block size, pop per-chunk
    results in order.  The engines must share one configuration and outlive this object."""

    def __init__(self, engines, nfft, overlap, device=0):
        self.lib = load()
        self.engines = list(engines)
        e0 = self.engines[0]
        self.D, self.M, self.max_sym = e0.D, e0.M, e0.max_sym
        arr = (_P * len(self.engines))(*[e._h for e in self.engines])
        self._h = _P()
        rc = self.lib.pcs_ingest_create(arr, len(self.engines), int(nfft), int(overlap), self.D, self.M, int(device),
                                        C.byref(self._h))
        if rc != 0:
            raise NativeError(rc, self.lib.pcs_last_error().decode())
        _live.add(self)
        self._E = np.empty((self.D, self.M), dtype=np.float32)
        self._sym = np.empty(self.max_sym, dtype=np.int32)
        self._centre = np.empty(self.max_sym, dtype=np.int32)
        self._mag = np.empty(self.max_sym, dtype=np.float32)
        self._sig = np.empty(8192, dtype=np.complex64)
        self._noise = np.empty(8192, dtype=np.complex64)

    def _check(self, rc):
        if rc != 0:
            raise NativeError(rc, self.lib.pcs_last_error().decode())

    def push(self, samples):
        """Append samples; returns the number of chunks this call completed and submitted."""
        samples = np.ascontiguousarray(samples, dtype=np.complex64)
        n = C.c_int32(0)
        self._check(self.lib.pcs_ingest_push(self._h, _ptr(samples), len(samples), C.byref(n)))
        return n.value

    def pop(self, block=True):
        """Next chunk's results ``(res, E, sym, centre, mag, sig_win, noise_win)`` (copies), or None if none is ready."""
        res, ready = Result(), C.c_int32(0)
        self._check(self.lib.pcs_ingest_pop(self._h, int(bool(block)), C.byref(res), _ptr(self._E), _ptr(self._sym),
                                            _ptr(self._centre), _ptr(self._mag), _ptr(self._sig), _ptr(self._noise),
                                            C.byref(ready)))
        if not ready.value:
            return None
        n, w = res.n_sym, max(res.sig_len, 0)
        return (res, self._E.copy(), self._sym[:n].copy(), self._centre[:n].copy(), self._mag[:n].copy(),
                self._sig[:w].copy(), self._noise[:w].copy())

    def pending(self):
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.pcs_ingest_pending(self._h, C.byref(a), C.byref(b)))
        return a.value - b.value

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._h = None
            self.lib.pcs_ingest_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """Thin object wrapper over one ``pcs_handle`` (one CUDA stream, one pinned chunk buffer)."""

    def __init__(self, *, device, nfft, num_dopplers, element_offset, shifts, masks, window_width, sum_all_masks,
                 code_search_mask_offset, samples_per_sym, path=PATH_AUTO, log2_block=0, snr_window=5, use_graph=True, groups_per_cta=0, xb_smem=False,
                 search_form=0, items_per_cta=0, warps20=False, fb_tree=True):
        self.lib = load()
        shifts = np.ascontiguousarray(shifts, dtype=np.int32)
        masks = np.ascontiguousarray(masks, dtype=np.complex64)
        if masks.ndim != 2 or masks.shape[1] != nfft:
            raise ValueError(f"masks must have shape (num_masks, {nfft}), got {masks.shape}")
        if shifts.shape != (num_dopplers + element_offset,):
            raise ValueError("shifts must have num_dopplers + element_offset entries")
        cfg = Config(ABI_VERSION, device, nfft, num_dopplers, element_offset, masks.shape[0], window_width,
                     int(bool(sum_all_masks)), code_search_mask_offset, samples_per_sym, path, log2_block, snr_window)
        cfg.reserved[0] = (0 if use_graph else 1) | (2 if warps20 else 0) | (0 if fb_tree else 4)   # bit 0: no CUDA graph; bit 1: 20-warp build; bit 2: factorised bank without shared partial sums
        cfg.reserved[1] = int(groups_per_cta) | (int(items_per_cta) << 8)    # tuning knobs of the 256-point search kernels
        # form of the 256-point search: 0 shifted filters (default), 1 rotate + block spectrum in shared memory, 2 rotate
        cfg.reserved[2] = 1 if xb_smem else int(search_form)
        self._h = _P()
        self.nfft, self.D, self.M = nfft, num_dopplers + element_offset, masks.shape[0]
        self._check(self.lib.pcs_create(C.byref(cfg), _ptr(shifts), _ptr(masks), C.byref(self._h)))
        _live.add(self)
        self.max_sym = self.lib.pcs_max_symbols(self._h)
        buf = (C.c_float * (2 * nfft)).from_address(self.lib.pcs_host_buffer(self._h))
        self.host_buffer = np.frombuffer(buf, dtype=np.complex64)
        self._E = np.empty((self.D, self.M), dtype=np.float32)
        self._sym = np.empty(self.max_sym, dtype=np.int32)
        self._centre = np.empty(self.max_sym, dtype=np.int32)
        self._mag = np.empty(self.max_sym, dtype=np.float32)

    def _check(self, rc):
        if rc != 0:
            raise NativeError(rc, self.lib.pcs_last_error().decode())

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self.host_buffer = None
            self._slots = {}
            self._h = None
            self.lib.pcs_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- per chunk ------------------------------------------------------------------------------
    def upload(self):
        self._check(self.lib.pcs_upload(self._h))

    def upload_thresholded(self, scale):
        """``__thresholdInput`` + ``uploadToGPU`` on the device (pcs_upload_thresholded): the pinned chunk ends up clipped
        in place; returns (indices clipped by the second pass, the two float32 clip levels)."""
        if getattr(self, "_clip_idx", None) is None:
            self._clip_idx = np.empty(self.nfft, dtype=np.int64)
            self._clip_thr = np.empty(2, dtype=np.float32)
        n = C.c_int32(0)
        self._check(self.lib.pcs_upload_thresholded(self._h, float(scale), _ptr(self._clip_idx), self.nfft, C.byref(n),
                                                    _ptr(self._clip_thr)))
        return self._clip_idx[:n.value].copy(), self._clip_thr.copy()

    def set_host_source(self, ptr):
        """The next ``upload`` / ``chunk_to_bits`` copies complex64[nfft] from this page-locked address (one shot)."""
        self._check(self.lib.pcs_set_host_source(self._h, _P(ptr)))

    def upload_device(self, dev_ptr):
        self._check(self.lib.pcs_upload_device(self._h, _P(dev_ptr)))

    def heterodyne(self, a, b=0.0, c=0.0):
        """De-chirp the uploaded chunk: x[n] * exp(j (a n^2 + b n + c)) becomes the chunk searched / demodulated next."""
        self._check(self.lib.pcs_heterodyne(self._h, float(a), float(b), float(c)))

    def chunk(self):
        x = np.empty(self.nfft, dtype=np.complex64)
        self._check(self.lib.pcs_get_chunk(self._h, _ptr(x)))
        return x

    def search(self):
        res = Result()
        self._check(self.lib.pcs_search(self._h, C.byref(res), _ptr(self._E)))
        return res, self._E

    def demod(self, shift=-1):
        res = Result()
        self._check(self.lib.pcs_demod(self._h, int(shift), C.byref(res), _ptr(self._sym), _ptr(self._centre),
                                       _ptr(self._mag)))
        n = res.n_sym
        return res, self._sym[:n], self._centre[:n], self._mag[:n]

    def process(self):
        res = Result()
        self._check(self.lib.pcs_process(self._h, C.byref(res), _ptr(self._E), _ptr(self._sym), _ptr(self._centre),
                                         _ptr(self._mag)))
        n = res.n_sym
        return res, self._E, self._sym[:n], self._centre[:n], self._mag[:n]

    def enqueue_device(self, dev_ptr):
        self._check(self.lib.pcs_enqueue_device(self._h, _P(dev_ptr)))

    def fetch(self):
        res = Result()
        self._check(self.lib.pcs_fetch(self._h, C.byref(res), _ptr(self._E), _ptr(self._sym), _ptr(self._centre),
                                       _ptr(self._mag)))
        n = res.n_sym
        return res, self._E, self._sym[:n], self._centre[:n], self._mag[:n]

    def snr_means(self, shifts):
        """(mean |X| signal window, mean |X| noise window) or None when the window geometry needs the general path."""
        a, b, ok = C.c_float(0), C.c_float(0), C.c_int32(0)
        self._check(self.lib.pcs_snr_means(self._h, _ptr(shifts), C.byref(a), C.byref(b), C.byref(ok)))
        return (np.float32(a.value), np.float32(b.value)) if ok.value else None

    def chunk_to_bits(self, stitcher, shifts, clipped):
        """upload + search + demod + SNR means + bit post-processing of the chunk in the pinned buffer, one native call.
        Returns (res, E, sym, centre, mag, snr_means or None, bits, centres, trust)."""
        n = self.max_sym
        if getattr(self, "_bits", None) is None:
            self._bits, self._cen8, self._tr8 = (np.empty(n, dtype=np.uint8) for _ in range(3))
        clipped = np.ascontiguousarray(clipped, dtype=np.int64)
        res, a, b, ok, n_out = Result(), C.c_float(0), C.c_float(0), C.c_int32(0), C.c_int32(0)
        rc = self.lib.pcs_chunk_to_bits(self._h, stitcher._h, _ptr(shifts), _ptr(clipped), len(clipped), C.byref(res),
                                        _ptr(self._E), _ptr(self._sym), _ptr(self._centre), _ptr(self._mag), C.byref(a),
                                        C.byref(b), C.byref(ok), _ptr(self._bits), _ptr(self._cen8), _ptr(self._tr8),
                                        C.byref(n_out))
        err = None
        if rc != 0 and n_out.value == -1:
            # the device part succeeded and the stitcher refused the symbol table (IndexError = what the reference's
            # np.where(...)[0][0] raises from demodulate()); handed back so that the caller can raise it there
            msg = self.lib.pcs_last_error().decode()
            err = IndexError(msg) if rc == -4 else NativeError(rc, msg)
        else:
            self._check(rc)
        k, ns = max(n_out.value, 0), res.n_sym
        means = (np.float32(a.value), np.float32(b.value)) if ok.value else None
        return (res, self._E, self._sym[:ns], self._centre[:ns], self._mag[:ns], means,
                self._bits[:k].copy(), self._cen8[:k].copy(), self._tr8[:k].copy(), err)

    def snr_windows(self, res):
        sig = np.empty(max(res.sig_len, 0), dtype=np.complex64)
        noise = np.empty(max(res.noise_len, 0), dtype=np.complex64)
        self._check(self.lib.pcs_snr_windows(self._h, _ptr(sig), _ptr(noise)))
        return sig, noise

    # -- inspection -----------------------------------------------------------------------------
    def spectrum(self):
        X = np.empty(self.nfft, dtype=np.complex64)
        self._check(self.lib.pcs_get_spectrum(self._h, _ptr(X)))
        return X

    def peaks(self):
        v = np.empty((self.D, self.M), dtype=np.float32)
        o = np.empty((self.D, self.M), dtype=np.int32)
        self._check(self.lib.pcs_get_peaks(self._h, _ptr(v), _ptr(o)))
        return v, o

    def demod_surface(self, shift):
        y = np.empty((self.M, self.nfft), dtype=np.complex64)
        self._check(self.lib.pcs_get_demod_surface(self._h, int(shift), _ptr(y)))
        return y

    def demod_magnitudes(self):
        ymag = np.empty((self.M, self.nfft), dtype=np.float32)
        p = np.empty(self.nfft, dtype=np.float32)
        self._check(self.lib.pcs_get_demod_magnitudes(self._h, _ptr(ymag), _ptr(p)))
        return ymag, p

    def plan(self):
        info = PlanInfo()
        self._check(self.lib.pcs_get_plan(self._h, C.byref(info)))
        return {n: getattr(info, n) for n, _ in PlanInfo._fields_}

    # -- bin sharding ----------------------------------------------------------------------------
    def bank_factor(self):
        """(form, S, J, R) of the factorised filter bank the generic search uses (``pcs_get_bank_factor``): form 0 = not
        factorised, 1 = general, 2 = complete binary bank, 3 = complete binary bank with shared partial sums."""
        out = np.zeros(4, dtype=np.int32)
        self._check(load().pcs_get_bank_factor(self._h, _ptr(out)))
        return tuple(int(v) for v in out)

    def set_bin_range(self, lo, hi):
        self._check(self.lib.pcs_set_bin_range(self._h, int(lo), int(hi)))

    def shard_buffers(self):
        """Device pointers of the [D*M] energy / peak-value / peak-offset tables."""
        e, v, o = _P(), _P(), _P()
        self._check(self.lib.pcs_shard_buffers(self._h, C.byref(e), C.byref(v), C.byref(o)))
        return e.value, v.value, o.value

    def enqueue_search_local(self):
        self._check(self.lib.pcs_enqueue_search_local(self._h))

    def enqueue_estimate_and_demod(self, with_demod=True):
        self._check(self.lib.pcs_enqueue_estimate_and_demod(self._h, int(bool(with_demod))))

    # -- bin-sharded streaming search over NVLink peer memory (pcs_shard_*) ---------------------------
    def shard_init(self, rank, world, ring=0):
        """Allocate the exchange region / chunk ring; returns its 64-byte CUDA IPC handle."""
        buf = C.create_string_buffer(64)
        self._check(self.lib.pcs_shard_init(self._h, int(rank), int(world), int(ring), buf))
        self._shard_world = int(world)
        self._slots = {}
        return buf.raw

    def shard_attach(self, handles):
        blob = b"".join(handles)
        if len(blob) != 64 * self._shard_world:
            raise ValueError("need one 64-byte IPC handle per rank")
        self._check(self.lib.pcs_shard_attach(self._h, C.c_char_p(blob)))

    def shard_info(self):
        v = [C.c_int32(0) for _ in range(5)]
        self._check(self.lib.pcs_shard_info(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("ring", "lanes", "bin_lo", "bin_hi", "stages"), (x.value for x in v)))

    def shard_host_slot(self, seq):
        """Ingest rank: NumPy view (complex64[nfft]) of the pinned slot chunk ``seq`` is read from."""
        p = _P()
        self._check(self.lib.pcs_shard_host_slot(self._h, int(seq), C.byref(p)))
        view = self._slots.get(p.value)
        if view is None:
            buf = (C.c_float * (2 * self.nfft)).from_address(p.value)
            view = self._slots[p.value] = np.frombuffer(buf, dtype=np.complex64)
        return view

    def shard_submit(self, seq, kind, src=None):
        """``src``: device pointer (int), host ndarray (complex64[nfft], contiguous) or None."""
        if isinstance(src, np.ndarray):
            if src.dtype != np.complex64 or not src.flags.c_contiguous or src.size != self.nfft:
                raise ValueError("host chunk must be a contiguous complex64[nfft] array")
            src = _ptr(src)
        self._check(self.lib.pcs_shard_submit(self._h, int(seq), int(kind), _P(src) if src else None))

    def shard_fetch(self, seq):
        """Owner: ``(res, E, sym, centre, mag, snr_means or None)`` of chunk ``seq`` (views into the engine's buffers)."""
        res, a, b, ok = Result(), C.c_float(0), C.c_float(0), C.c_int32(0)
        self._check(self.lib.pcs_shard_fetch(self._h, int(seq), C.byref(res), _ptr(self._E), _ptr(self._sym), _ptr(self._centre),
                                             _ptr(self._mag), C.byref(a), C.byref(b), C.byref(ok)))
        n = res.n_sym
        means = (np.float32(a.value), np.float32(b.value)) if ok.value else None
        return res, self._E, self._sym[:n], self._centre[:n], self._mag[:n], means

    def shard_sync(self):
        self._check(self.lib.pcs_shard_sync(self._h))

    def shard_trace(self):
        """(first chunk, float32[n, 3]) timeline in ms: block spectra start, search end, tail end (PCS_SHARD_TRACE=1)."""
        first, n = C.c_int64(0), C.c_int32(0)
        out = np.zeros((64, 3), dtype=np.float32)
        self._check(self.lib.pcs_shard_trace(self._h, C.byref(first), C.byref(n), _ptr(out)))
        return first.value, out[:n.value].copy()

    def shard_streams(self):
        """(lane 0, lane 1, copy, tail) CUDA streams of the engine as integers."""
        out = (C.c_uint64 * 4)()
        self._check(self.lib.pcs_shard_streams(self._h, out))
        return tuple(int(v) for v in out)

    def set_stream(self, stream_ptr):
        self._check(self.lib.pcs_set_stream(self._h, int(stream_ptr)))

    STAGES = ("spectrum", "search", "estimate", "demod_surface", "timing_symbols", "reduce", "block_spectra")

    def set_profiling(self, enable=True):
        self._check(self.lib.pcs_set_profiling(self._h, int(bool(enable))))

    def profile(self):
        """{stage: (total_ms, count)} accumulated since profiling was switched on."""
        ms = np.zeros(len(self.STAGES), dtype=np.float64)
        cnt = np.zeros(len(self.STAGES), dtype=np.int64)
        self._check(self.lib.pcs_get_profile(self._h, _ptr(ms), _ptr(cnt)))
        return {n: (float(ms[i]), int(cnt[i])) for i, n in enumerate(self.STAGES)}

    @property
    def launch_count(self):
        return self.lib.pcs_launch_count(self._h)

    @property
    def stream(self):
        return self.lib.pcs_stream(self._h)
