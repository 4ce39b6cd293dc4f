"""The native (C++) bit post-processing against the oracle's host functions -- which are themselves pinned to the
reference's own Python (tests/test_oracle_golden.py).  Runs on the CPU: the stitcher has no device code."""
import numpy as np
import pytest

from oracle import oracle as O
from pycusdr_b200 import _native
from pycusdr_b200.protocol.benchmark.bench_protocols import _nrzs_lut

N, OVL, OO, ERR_THR, MATCH_THR = 8192, 1024, 20, 1000, 10


def _oracle_chunk(state, sym, centres, mag, clipped, spSym, bitLUT, symbolLUT):
    trust = O.trust_from_magnitudes(mag, len(sym))
    bits, err = O.extract_bits(sym, bitLUT, symbolLUT)
    cW, bW, tW = O.check_symbol_overlap(state, len(err), centres, bits, trust, N, OVL // 2, OO, ERR_THR, MATCH_THR)
    tW = O.tag_clipped_peaks(tW, cW, clipped, spSym, N)
    return bW.astype(np.uint8), cW.astype(np.uint8), tW.astype(np.uint8)


def _stream(rng, n_chunks, M, sps, slip_at=(), noisy_at=()):
    """Symbol decisions of consecutive overlapping chunks of one long random symbol sequence."""
    total = rng.randint(0, M, size=n_chunks * (N // sps) + 64).astype(np.int32)
    step_syms = (N - OVL) // sps
    out = []
    for c in range(n_chunks):
        off = rng.uniform(0, sps)
        S = int(N / sps)
        k0 = c * step_syms + (1 if c in slip_at else 0) - (1 if (c + 100) in slip_at else 0)
        sym = total[k0:k0 + S].copy()
        if c in noisy_at:
            flip = rng.rand(S) < 0.3
            sym[flip] = rng.randint(0, M, size=int(flip.sum()))
        if c == 2:
            sym[5] = -1                                         # all-zero window (kern:104-105 initial values)
        centres = (np.arange(S) * sps + off + rng.randint(-2, 3, size=S)).astype(np.int32)
        mag = (rng.rand(S) * 10 ** rng.uniform(-3, 3)).astype(np.float32)
        out.append((sym, centres, mag))
    return out


@pytest.mark.parametrize("kind", ["bitlut", "nrzs"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_native_stitcher_matches_oracle(kind, seed):
    rng = np.random.RandomState(seed)
    if kind == "bitlut":
        M, bitLUT, symbolLUT = 8, np.array([0, 0, 1, 1, 0, 0, 1, 1]), []
    else:
        M, bitLUT, symbolLUT = 16, None, _nrzs_lut(5)
    st = _native.Stitcher(nfft=N, overlap=OVL, overlap_offset=OO, error_threshold=ERR_THR, match_threshold=MATCH_THR,
                          bit_lut=bitLUT, symbol_lut=symbolLUT)
    state = O.OverlapState()
    chunks = _stream(rng, 14, M, 16, slip_at=(3, 7, 108, 111), noisy_at=(5, 9))
    adjusted = 0
    for c, (sym, centres, mag) in enumerate(chunks):
        clipped = np.sort(rng.choice(N, size=rng.randint(0, 4), replace=False)) if c % 3 == 1 else np.array([], dtype=np.int64)
        if c == 4:
            clipped = np.array([3, 10, N - 2])                 # slices that start below zero / end past the chunk
        spSym = 16.0 + rng.uniform(-0.4, 0.4)
        want = _oracle_chunk(state, sym, centres, mag, clipped, spSym, bitLUT, symbolLUT)
        got = st(sym, centres, mag, clipped, spSym)
        for g, w, name in zip(got, want, ("bits", "centres", "trust")):
            np.testing.assert_array_equal(g, w, err_msg=f"chunk {c}: {name}")
        adjusted += int(len(want[0]) != np.sum((centres >= OVL // 2) & (np.arange(len(centres)) < np.argmax(centres > N - OVL // 2))))
    assert adjusted >= 1, "the +-1 symbol realignment branch was never taken: the test stream is too tame"


def test_native_stitcher_errors_mirror_the_reference():
    st = _native.Stitcher(nfft=N, overlap=OVL, overlap_offset=OO, error_threshold=ERR_THR, match_threshold=MATCH_THR,
                          bit_lut=np.array([0, 1]), symbol_lut=[])
    sym = np.zeros(16, np.int32)
    with pytest.raises(IndexError):                            # no centre beyond Nfft - overlap/2
        st(sym, np.arange(16, dtype=np.int32) * 16, np.ones(16, np.float32), [], 16.0)
    with pytest.raises(_native.NativeError):                   # symbol outside the table
        st(sym + 5, np.arange(16, dtype=np.int32) * 300, np.ones(16, np.float32), [], 16.0)
    with pytest.raises(NotImplementedError):
        _native.Stitcher(nfft=N, overlap=OVL, overlap_offset=OO, error_threshold=ERR_THR, match_threshold=MATCH_THR,
                         bit_lut=None, symbol_lut=np.zeros((4, 3)))


@pytest.mark.parametrize("seed", [0, 1])
def test_sync_search_equals_numpy_convolve(seed):
    """decoder.py:96-104 on the handed-over bit stream: native integer search vs np.convolve."""
    import time
    rng = np.random.RandomState(seed)
    header = rng.randint(0, 2, 128)
    mask = np.flipud(header * 2 - 1)                      # protocol.get_mask() of the benchmark protocols
    bits = rng.randint(0, 2, 20000).astype(np.uint8)
    for pos in (0, 777, 9000, 20000 - 128):               # plant headers with a few bit errors
        h = header.copy()
        h[rng.choice(128, size=rng.randint(0, 20), replace=False)] ^= 1
        bits[pos:pos + 128] = h
    thr = int(np.sum(header)) - 27                        # numOnesHeader - headerTol (bench_base.py:62-64)
    t0 = time.perf_counter()
    score = np.convolve(bits.astype(np.float64), mask)
    want = np.where(score >= thr)[0]
    t_np = time.perf_counter() - t0
    t0 = time.perf_counter()
    idx, sc = _native.sync_search(bits, mask, thr)
    t_nat = time.perf_counter() - t0
    np.testing.assert_array_equal(idx, want)
    np.testing.assert_array_equal(sc, score[want].astype(np.int32))
    assert set((want - 127).tolist()) >= {0, 777, 9000, 20000 - 128}
    assert len(_native.sync_search(np.zeros(0, np.uint8), mask, thr)[0]) == 0
    # every candidate when the threshold is low enough to overflow the first output buffer
    idx2, _ = _native.sync_search(bits, mask, -10)
    np.testing.assert_array_equal(idx2, np.where(score >= -10)[0])
    print(f"np.convolve {t_np * 1e3:.2f} ms, native {t_nat * 1e3:.2f} ms")


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_injected_carry_is_the_references_state(seed):
    """pcs_stitch_set_state / get_state speak the reference's own state variables (poswinP, posSymEnd, dem_base:977-979):
    a stitcher primed with the oracle's state processes the next chunk exactly like the oracle, whatever it processed
    before, and ends up holding the oracle's new state."""
    rng = np.random.RandomState(100 + seed)
    M, bitLUT = 8, np.array([0, 0, 1, 1, 0, 0, 1, 1])
    mk = lambda: _native.Stitcher(nfft=N, overlap=OVL, overlap_offset=OO, error_threshold=ERR_THR,      # noqa: E731
                                  match_threshold=MATCH_THR, bit_lut=bitLUT, symbol_lut=[])
    chunks = _stream(rng, 8, M, 16, slip_at=(2, 5, 104), noisy_at=(3,))
    state = O.OverlapState()
    stray = mk()
    for c, (sym, centres, mag) in enumerate(chunks):
        if c == 2:
            continue                                           # (the chunk with a -1 symbol is covered above)
        fresh = mk()
        if c % 2:
            stray(*chunks[(c + 3) % 8 if (c + 3) % 8 != 2 else 4], [], 16.0)    # leave some unrelated carry behind
            fresh = stray
        pos = np.asarray(state.poswinP, dtype=np.uint8)
        end = np.asarray(state.posSymEnd if state.posSymEnd is not None else [], dtype=np.uint8)
        token = len(pos).to_bytes(4, "little") + len(end).to_bytes(4, "little") + pos.tobytes() + end.tobytes()
        fresh.set_state(token)
        assert fresh.get_state() == token
        want = _oracle_chunk(state, sym, centres, mag, np.array([], dtype=np.int64), 16.0, bitLUT, [])
        got = fresh(sym, centres, mag, [], 16.0)
        for g, w in zip(got, want):
            np.testing.assert_array_equal(g, w)
        new = fresh.get_state()
        a, b = int.from_bytes(new[:4], "little"), int.from_bytes(new[4:8], "little")
        np.testing.assert_array_equal(np.frombuffer(new, np.uint8, a, 8), np.asarray(state.poswinP, dtype=np.uint8))
        np.testing.assert_array_equal(np.frombuffer(new, np.uint8, b, 8 + a), np.asarray(state.posSymEnd, dtype=np.uint8))


def test_carry_of_a_large_overlap_fits_whatever_the_geometry():
    """overlap 2^13 at 4 samples per symbol carries ~2 K bits; the token is sized from the stitcher's own lengths
    (and state_capacity bounds it for fixed-size transports)."""
    nfft, ovl, sps = 2 ** 15, 2 ** 13, 4
    st = _native.Stitcher(nfft=nfft, overlap=ovl, overlap_offset=OO, error_threshold=ERR_THR, match_threshold=MATCH_THR,
                          bit_lut=np.array([0, 0, 1, 1, 0, 0, 1, 1]), symbol_lut=[])
    n = nfft // sps
    rng = np.random.RandomState(3)
    sym = rng.randint(0, 8, n).astype(np.int32)
    centres = (np.arange(n) * sps + 1).astype(np.int32)
    mag = rng.rand(n).astype(np.float32)
    st(sym, centres, mag, [], float(sps))
    token = st.get_state()
    a, b = int.from_bytes(token[:4], "little"), int.from_bytes(token[4:8], "little")
    assert a > 1020 and b == OO + 1 and len(token) == 8 + a + b
    assert len(token) <= _native.Stitcher.state_capacity(ovl, OO, sps // 2, 7)
    other = _native.Stitcher(nfft=nfft, overlap=ovl, overlap_offset=OO, error_threshold=ERR_THR, match_threshold=MATCH_THR,
                             bit_lut=np.array([0, 0, 1, 1, 0, 0, 1, 1]), symbol_lut=[])
    other.set_state(token)
    assert other.get_state() == token


def _circular_xcorr_exact(a, b, n):
    """c[k] = sum_j a_pad[(j + k) % n] * b_pad[j] with np.correlate on the doubled ring: exact integers."""
    ap = np.r_[a, np.zeros(n - len(a))].astype(np.float64)
    bp = np.asarray(b, dtype=np.float64)
    ring = np.r_[ap, ap[:max(len(bp) - 1, 0)]]
    return np.rint(np.correlate(ring, bp, mode="valid")[:n]).astype(np.int32)


@pytest.mark.parametrize("n_slave,n_master", [(37, 20), (64, 64), (1000, 400), (5000, 3000), (20000, 2500), (9000, 12000)])
def test_bit_xcorr_is_the_exact_circular_correlation(n_slave, n_master):
    """pcs_bit_xcorr (softCombiner.py:697-706 + lib/customXCorr.py:5-17): exact integers, equal to np.correlate on the
    ring and to the reference's FFT formula up to its floating-point rounding."""
    rng = np.random.RandomState(n_slave + n_master)
    slave = rng.randint(0, 2, n_slave).astype(np.uint8)
    master = rng.randint(0, 2, n_master).astype(np.uint8)
    m = master[:n_slave]                                   # the reference passes bitsM[:n]
    got = _native.bit_xcorr(slave, m)
    n = int(2 ** np.ceil(np.log2(n_slave)))
    assert got.dtype == np.int32 and len(got) == n
    np.testing.assert_array_equal(got, _circular_xcorr_exact(slave, m, n))
    # the reference's own formula: np.abs(ifft(fft(bitsX, N) * conj(fft(bitsM[:n], N)), N)) with bitsX zero padded to N
    bitsX = np.r_[slave, np.zeros(n - n_slave)]
    ref = np.abs(np.fft.ifft(np.fft.fft(bitsX, n) * np.conj(np.fft.fft(m, n)), n))
    assert np.max(np.abs(got - ref)) < 1e-6
    # odd ring sizes take the plain path
    got_odd = _native.bit_xcorr(slave, m, n=n_slave + 3)
    np.testing.assert_array_equal(got_odd, _circular_xcorr_exact(slave, m, n_slave + 3))


def test_align_bits_finds_the_masters_position_like_the_reference():
    """softCombiner.correlate's acceptance test (softCombiner.py:708-722): the master's new bits sit somewhere in the
    slave's history; idx[0] is where, and val[0] must clear mean + varMultiplier * std of val[2:]."""
    rng = np.random.RandomState(4)
    master = rng.randint(0, 2, 3000).astype(np.uint8)
    for off, ber in ((0, 0.0), (1234, 0.02), (6000, 0.08)):
        slave = rng.randint(0, 2, 10000).astype(np.uint8)
        L = min(len(master), len(slave) - off)
        seg = master[:L].copy()
        seg[rng.rand(L) < ber] ^= 1
        slave[off:off + L] = seg
        pos, ok, idx, val, cond = _native.align_bits(slave, master, 3.0)
        # the reference's statements on its own (floating point) correlation
        n = len(slave)
        N = int(2 ** np.ceil(np.log2(n)))
        x = np.abs(np.fft.ifft(np.fft.fft(np.r_[slave, np.zeros(N - n)], N) * np.conj(np.fft.fft(master[:n], N)), N))
        ridx, rval = np.empty(15, int), np.empty(15)
        for i in range(15):
            ridx[i] = np.argmax(x)
            rval[i] = x[ridx[i]]
            x[ridx[i]] = 0
        rcond = np.mean(rval[2:]) + 3.0 * np.std(rval[2:])
        assert (pos, ok) == (off, True) == (int(ridx[0]), bool(rval[0] > rcond))
        np.testing.assert_allclose(val, rval, atol=1e-6)
        assert abs(cond - rcond) < 1e-6
    # unrelated streams are rejected by both
    pos, ok, idx, val, cond = _native.align_bits(rng.randint(0, 2, 8000), rng.randint(0, 2, 3000), 5.0)
    assert not ok
