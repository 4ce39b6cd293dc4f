// Host-side bit post-processing of one chunk in C++: what the reference does in NumPy after the
// three D2H copies of cudaFindCentres -- extractBits / extractBitsNRZs (dem_base:1012-1051), checkSymbolOverlap
// (dem_base:863-988), clipped-peak tagging of the trust (dem_base:817-837) and the uint8 output casts (dem_base:859).
// No device code; it lives in the same shared library so the Python mirror makes one call per chunk.
// Python slice semantics (negative starts, clamping, empty results) are reproduced exactly because the reference's
// stitching relies on them at the chunk edges.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pycusdr_b200.h"

int pcs_fail_msg(int code, const char* msg);   // pcs_api.cu

namespace {

struct Span {
    long lo, hi;   // [lo, hi), already clamped to the array
    long len() const { return hi > lo ? hi - lo : 0; }
};
const long NONE = INT64_MIN;

// a[start:stop] on an array of n elements, Python rules (step 1)
Span pyslice(long n, long start, long stop) {
    long lo = start == NONE ? 0 : (start < 0 ? start + n : start);
    long hi = stop == NONE ? n : (stop < 0 ? stop + n : stop);
    if (lo < 0) lo = 0;
    if (lo > n) lo = n;
    if (hi < 0) hi = 0;
    if (hi > n) hi = n;
    if (hi < lo) hi = lo;
    return {lo, hi};
}

// number of equal elements of a[sa] and b[sb], or -1 when the shapes differ (NumPy: comparison is "False")
long matches(const uint8_t* a, Span sa, const uint8_t* b, Span sb) {
    if (sa.len() != sb.len()) return -1;
    long c = 0;
    for (long i = 0; i < sa.len(); ++i) c += a[sa.lo + i] == b[sb.lo + i];
    return c;
}

}  // namespace

struct pcs_stitcher {
    pcs_stitch_config cfg;
    std::vector<uint8_t> bit_lut;       // [M]
    std::vector<int32_t> symbol_lut;    // [M][2][K]
    std::vector<uint8_t> poswinP, posSymEnd, bits, near;
    bool have_prev = false;
};

extern "C" {

int pcs_stitch_create(const pcs_stitch_config* cfg, const uint8_t* bit_lut, const int32_t* symbol_lut, pcs_stitcher** out) {
    if (!cfg || !out) return pcs_fail_msg(PCS_ERR_INVALID, "null argument");
    if (cfg->num_symbols < 1 || cfg->nfft < 1) return pcs_fail_msg(PCS_ERR_INVALID, "bad stitcher geometry");
    if (!bit_lut && !(symbol_lut && cfg->lut_k > 0))
        return pcs_fail_msg(PCS_ERR_INVALID, "neither a bit LUT nor a 3-D symbol LUT: extractBitsOld is not defined by the "
                                             "reference either (dem_base:1017)");
    pcs_stitcher* s = new pcs_stitcher();
    s->cfg = *cfg;
    if (bit_lut) s->bit_lut.assign(bit_lut, bit_lut + cfg->num_symbols);
    else s->symbol_lut.assign(symbol_lut, symbol_lut + (size_t)cfg->num_symbols * 2 * cfg->lut_k);
    *out = s;
    return PCS_OK;
}

int pcs_stitch_reset(pcs_stitcher* s) {
    if (!s) return pcs_fail_msg(PCS_ERR_INVALID, "null stitcher");
    s->poswinP.clear();
    s->posSymEnd.clear();
    s->have_prev = false;
    return PCS_OK;
}

int pcs_stitch_destroy(pcs_stitcher* s) {
    delete s;
    return PCS_OK;
}

// The carry between consecutive chunks (dem_base:977-979: poswinP = the bits behind the window, posSymEnd = the last
// overlap_offset + 1 bits of the window).  It is a function of its own chunk's symbols alone, which is what lets chunks be
// post-processed on different ranks: the owner of chunk k hands this to the owner of chunk k + 1.
int pcs_stitch_get_state(const pcs_stitcher* s, uint8_t* buf, int32_t cap, int32_t* n_poswin, int32_t* n_posend) {
    if (!s || !n_poswin || !n_posend) return pcs_fail_msg(PCS_ERR_INVALID, "null argument");
    const size_t a = s->poswinP.size(), b = s->posSymEnd.size();
    *n_poswin = (int32_t)a;
    *n_posend = (int32_t)b;
    if (a + b > (size_t)std::max(cap, 0) || (a + b > 0 && !buf)) return pcs_fail_msg(PCS_ERR_INVALID, "state buffer too small");
    if (a) memcpy(buf, s->poswinP.data(), a);
    if (b) memcpy(buf + a, s->posSymEnd.data(), b);
    return PCS_OK;
}

int pcs_stitch_set_state(pcs_stitcher* s, const uint8_t* buf, int32_t n_poswin, int32_t n_posend) {
    if (!s || n_poswin < 0 || n_posend < 0 || (n_poswin + n_posend > 0 && !buf)) return pcs_fail_msg(PCS_ERR_INVALID, "bad argument");
    s->poswinP.assign(buf, buf + n_poswin);
    s->posSymEnd.assign(buf + n_poswin, buf + n_poswin + n_posend);
    s->have_prev = n_poswin > 0;
    return PCS_OK;
}

int pcs_stitch_chunk(pcs_stitcher* s, const int32_t* sym, const int32_t* centre, const float* mag, int32_t n_sym,
                     const int64_t* clipped, int32_t n_clipped, double sp_sym, uint8_t* bits_out, uint8_t* centres_out,
                     uint8_t* trust_out, int32_t* n_out) {
    if (!s || !sym || !centre || !mag || !bits_out || !centres_out || !trust_out || !n_out)
        return pcs_fail_msg(PCS_ERR_INVALID, "null argument");
    if (n_sym < 1) return pcs_fail_msg(PCS_ERR_INVALID, "no symbols");
    const pcs_stitch_config& c = s->cfg;
    const int M = c.num_symbols;
    // trust = the first n_sym raw bytes of the float magnitudes (dem_base:472,1005-1007)
    const uint8_t* trust = reinterpret_cast<const uint8_t*>(mag);
    // ---- extractBits / extractBitsNRZs ----
    std::vector<uint8_t>& bits = s->bits;
    long n_bits, n_err = 0;
    auto wrap = [M](int v) { return v < 0 ? v + M : v; };
    {   // every index inside [-M, M) (NumPy's negative indices included): one branch-free pass
        unsigned bad = 0;
        for (int i = 0; i < n_sym; ++i) bad |= (unsigned)(sym[i] + M) >= (unsigned)(2 * M);
        if (bad) return pcs_fail_msg(PCS_ERR_INVALID, "symbol index outside the look-up table");
    }
    if (!s->bit_lut.empty()) {
        n_bits = n_sym;
        bits.resize(n_bits);
        uint8_t lut2[128];                                   // lut2[v + M] = bit_lut[v mod M] for v in [-M, M)
        std::vector<uint8_t> big;
        uint8_t* l2 = lut2;
        if (2 * M > 128) { big.resize((size_t)2 * M); l2 = big.data(); }
        for (int v = 0; v < M; ++v) l2[v] = l2[v + M] = s->bit_lut[v];
        uint8_t* __restrict__ bo = bits.data();
        for (long i = 0; i < n_bits; ++i) bo[i] = l2[sym[i] + M];                       // dem_base:1020
    } else {
        const int K = c.lut_k;
        n_bits = n_sym - 1;
        bits.resize(n_bits);
        for (long i = 0; i < n_bits; ++i) {                                             // dem_base:1041-1051
            const int32_t* row = &s->symbol_lut[(size_t)wrap(sym[i]) * 2 * K];
            bool one = false, zero = false;
            for (int k = 0; k < K; ++k) {
                one |= row[k] == sym[i + 1];
                zero |= row[K + k] == sym[i + 1];
            }
            if (!(one || zero)) { ++n_err; one = false; }                               // SYMBOL_MISMATCHVAL = 0
            bits[i] = one;
        }
    }
    // ---- checkSymbolOverlap ----
    const int half = c.overlap / 2;                                                     // sigOverlapWin, dem_base:91
    long start = -1, end = -1;
    for (long i = 0; i < n_sym; ++i)
        if (centre[i] >= half) { start = i; break; }
    {   // first centre beyond nfft - half: near the END of the chunk, so look block by block (a max the compiler vectorises)
        const int thr = c.nfft - half;
        for (long i0 = 0; i0 < n_sym && end < 0; i0 += 256) {
            const long i1 = std::min<long>(i0 + 256, n_sym);
            int mx = INT32_MIN;
            for (long i = i0; i < i1; ++i) mx = std::max(mx, centre[i]);
            if (mx > thr)
                for (long i = i0; i < i1; ++i)
                    if (centre[i] > thr) { end = i; break; }
        }
    }
    if (start < 0 || end < 0)
        return pcs_fail_msg(PCS_ERR_STATE, "no symbol centre inside the overlap windows (the reference raises IndexError here)");
    const long oo = c.overlap_offset;
    const uint8_t* b = bits.data();
    if (n_err <= c.error_threshold && !s->poswinP.empty()) {
        const uint8_t* P = s->poswinP.data();
        const uint8_t* E = s->posSymEnd.data();
        const long nP = (long)s->poswinP.size(), nE = (long)s->posSymEnd.size();
        const Span win = pyslice(n_bits, start, end), pre = pyslice(n_bits, NONE, start);
        auto sub = [](Span outer, long a, long z) {      // outer[a:z]
            Span in = pyslice(outer.len(), a, z);
            return Span{outer.lo + in.lo, outer.lo + in.hi};
        };
        struct Pair { const uint8_t* a; Span sa; const uint8_t* b; Span sb; };
        const Pair pairs[6] = {
            {P, pyslice(nP, NONE, oo), b, sub(win, NONE, oo)},                      // pre
            {E, pyslice(nE, -oo, NONE), b, sub(pre, -oo, NONE)},                    // pos
            {P, pyslice(nP, NONE, oo), b, sub(win, 1, oo + 1)},                     // earlyPre
            {E, pyslice(nE, -oo - 1, -1), b, sub(pre, -oo, NONE)},                  // earlyPos
            {P, pyslice(nP, 1, oo + 1), b, sub(win, 0, oo)},                        // latePre
            {E, pyslice(nE, -oo, NONE), b, sub(pre, -oo - 1, -1)},                  // latePos
        };
        long n[6];
        for (int k = 0; k < 6; ++k) n[k] = matches(pairs[k].a, pairs[k].sa, pairs[k].b, pairs[k].sb);
        auto full = [&](int k) { return n[k] >= 0 && n[k] == pairs[k].sa.len(); };
        if (!(full(0) || full(1))) {
            long cnt[6];
            for (int k = 0; k < 6; ++k) cnt[k] = n[k] < 0 ? 0 : n[k];
            const long maxPre = std::max(cnt[0], std::max(cnt[2], cnt[4]));
            const long maxPos = std::max(cnt[1], std::max(cnt[3], cnt[5]));
            const long thr = c.match_threshold;
            if (thr < cnt[2] && cnt[2] == maxPre) {
                if (thr < cnt[3] && cnt[3] == maxPos) start += 1;                       // drop the first bit
            } else if (thr < cnt[4] && cnt[4] == maxPre) {
                if (thr < cnt[5] && cnt[5] == maxPos) start -= 1;                       // re-insert the last pre-window bit
            }
        }
    }
    const Span w = pyslice(n_bits, start, end);           // dataBits[start:end]
    const Span wc = pyslice(n_sym, start, end);           // centres / trust [start:end]
    const long n_win = w.len();
    // carry for the next chunk (dem_base:977-979)
    {
        const Span tail = pyslice(n_bits, end, NONE);
        s->poswinP.assign(b + tail.lo, b + tail.hi);
        const Span last = pyslice(n_win, -oo - 1, NONE);
        s->posSymEnd.assign(b + w.lo + last.lo, b + w.lo + last.hi);
    }
    // ---- outputs + clipped-peak tagging ----
    // the three arrays have the same length except in the NRZ-S path at the very end of the chunk; the reference then
    // fails on the boolean index -- here the shorter length is used
    const long n_ret = std::min(n_win, wc.len());
    if (n_ret > 0) {
        memcpy(bits_out, b + w.lo, (size_t)n_ret);
        memcpy(trust_out, trust + wc.lo, (size_t)n_ret);
        uint8_t* __restrict__ co = centres_out;
        const int32_t* __restrict__ ci = centre + wc.lo;
        for (long i = 0; i < n_ret; ++i) co[i] = (uint8_t)ci[i];                        // .astype(np.uint8), dem_base:859
    }
    if (n_clipped > 0) {
        const long N = c.nfft;
        const long span = 2 * (long)ceil(sp_sym);
        s->near.assign((size_t)N, 0);
        for (int k = 0; k < n_clipped; ++k) {
            const Span r = pyslice(N, (long)clipped[k] - span, (long)clipped[k] + span + 1);
            if (r.len() > 0) memset(s->near.data() + r.lo, 1, (size_t)r.len());
        }
        for (long i = 0; i < n_ret; ++i) {
            long ci = centre[wc.lo + i];
            if (ci < 0) ci += N;
            if (ci < 0 || ci >= N) return pcs_fail_msg(PCS_ERR_STATE, "symbol centre outside the chunk while tagging clipped peaks");
            if (s->near[ci]) trust_out[i] = (uint8_t)(int8_t)-2;
        }
    }
    *n_out = (int32_t)n_ret;
    return PCS_OK;
}

// clippedPeakI of __thresholdInput (dem_base:686-705): the clipped sample indices with every gap shorter than min_gap
// samples between two of them filled in.  idx ascending and unique (what the clip pass returns); out ascending.
int pcs_fill_gaps(const int64_t* idx, int32_t n, int32_t min_gap, int64_t* out, int32_t cap, int32_t* n_out) {
    if (!n_out || n < 0 || (n > 0 && !idx) || cap < 0 || (cap > 0 && !out)) return pcs_fail_msg(PCS_ERR_INVALID, "bad argument");
    int64_t k = 0;
    auto emit = [&](int64_t v) {
        if (k < cap) out[k] = v;
        ++k;
    };
    for (int32_t i = 0; i < n; ++i) {
        emit(idx[i]);
        if (i + 1 < n) {
            const int64_t step = idx[i + 1] - idx[i];
            if (step < 1) return pcs_fail_msg(PCS_ERR_INVALID, "indices must be ascending and unique");
            if (step > 1 && step < min_gap)
                for (int64_t v = idx[i] + 1; v < idx[i + 1]; ++v) emit(v);
        }
    }
    if (k > INT32_MAX) return pcs_fail_msg(PCS_ERR_INVALID, "too many indices");
    *n_out = (int32_t)k;       // may exceed cap: the list was truncated
    return PCS_OK;
}

// ---- decoder-side sync search ---------------------------------------------------------------
// What decoder.py:96-104 does with NumPy on the bit stream the demodulator hands over:
//   score = np.convolve(bits, mask)            (full convolution, mask = flipud(header * 2 - 1), protocol.get_mask())
//   idxCand = np.where(score >= numOnesHeader - headerTol)[0];   packetIdx = idxCand - len(mask) + 1
// Integer arithmetic, so the result is exactly NumPy's.  Returns the candidates (and their scores) in increasing order;
// *n_found is the total number of candidates even when it exceeds `cap`.
int pcs_sync_search(const uint8_t* bits, int64_t n, const int8_t* mask, int32_t m, int32_t threshold, int32_t* idx_out,
                    int32_t* score_out, int32_t cap, int32_t* n_found) {
    if (!bits || !mask || !n_found || n < 0 || m < 1) return pcs_fail_msg(PCS_ERR_INVALID, "bad argument");
    // score[i] = sum_j bits[j] * mask[i - j],  0 <= j < n, 0 <= i - j < m,  i in [0, n + m - 1)
    std::vector<int8_t> rev(mask, mask + m);                 // rev[k] = mask[m - 1 - k]: the header in +-1 form
    for (int k = 0; k < m / 2; ++k) std::swap(rev[k], rev[m - 1 - k]);
    int32_t found = 0;
    const int64_t total = n > 0 ? n + m - 1 : 0;
    auto emit = [&](int64_t i, int32_t sc) {
        if (sc >= threshold) {
            if (found < cap) {
                if (idx_out) idx_out[found] = (int32_t)i;
                if (score_out) score_out[found] = sc;
            }
            ++found;
        }
    };
    bool pm1 = true, binary = true;
    for (int k = 0; k < m; ++k) pm1 &= (rev[k] == 1 || rev[k] == -1);
    for (int64_t j = 0; j < n; ++j) binary &= bits[j] <= 1;
    if (pm1 && binary) {
        // +-1 header on a 0/1 stream: score = popcount(window & H) - popcount(window & ~H) on a multi-word shift register
        // (bit k of the register = window sample k, the newest sample enters at bit m - 1)
        const int W = (m + 63) / 64;
        std::vector<uint64_t> H(W, 0), nH(W, 0), reg(W, 0);
        for (int k = 0; k < m; ++k) (rev[k] > 0 ? H : nH)[k >> 6] |= 1ull << (k & 63);
        const int top = (m - 1) >> 6, topbit = (m - 1) & 63;
        for (int64_t i = 0; i < total; ++i) {
            for (int w = 0; w < W - 1; ++w) reg[w] = (reg[w] >> 1) | (reg[w + 1] << 63);
            reg[W - 1] >>= 1;
            if (i < n && bits[i]) reg[top] |= 1ull << topbit;
            int32_t sc = 0;
            for (int w = 0; w < W; ++w) sc += __builtin_popcountll(reg[w] & H[w]) - __builtin_popcountll(reg[w] & nH[w]);
            emit(i, sc);
        }
    } else {
        std::vector<uint8_t> pad((size_t)n + 2 * (size_t)(m - 1), 0);
        if (n > 0) memcpy(pad.data() + (m - 1), bits, (size_t)n);
        for (int64_t i = 0; i < total; ++i) {
            const uint8_t* w = pad.data() + i;                   // window bits[i - m + 1 .. i], zero padded
            int32_t sc = 0;
            for (int k = 0; k < m; ++k) sc += (int32_t)w[k] * (int32_t)rev[k];
            emit(i, sc);
        }
    }
    *n_found = found;
    return PCS_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Soft-combiner bit-stream alignment (softCombiner.py:697-722, lib/customXCorr.py:5-30).
// The reference zero-pads the slave's bits to N = 2^ceil(log2 n), and takes |IFFT(FFT(slave, N) . conj(FFT(master[:n], N)))|:
// the circular cross-correlation  c[k] = sum_j slave[(j + k) mod N] * master[j]  -- for 0/1 streams the number of positions
// where both hold a one -- then the 15 largest values by repeated arg-max.  Here c[k] is computed EXACTLY, in integers:
// 64 bit-rotated copies of the packed slave stream, then popcount(copy_s[(w + q) mod W] & master[w]) for k = 64 q + s.
// (The reference's values are these integers plus FFT rounding noise; where two shifts hold the same count its arg-max is
// decided by that noise, here the lowest index wins.)
// ---------------------------------------------------------------------------------------------------------------------
int pcs_bit_xcorr(const uint8_t* a, int64_t na, const uint8_t* b, int64_t nb, int64_t n, int32_t* out) {
    if (!a || !b || !out || na < 0 || nb < 0 || n < 1 || na > n || nb > n) return pcs_fail_msg(PCS_ERR_INVALID, "pcs_bit_xcorr: bad argument");
    if ((n & 63) != 0) {                                    // short or odd rings: plain sums
        std::vector<int64_t> ones;
        for (int64_t j = 0; j < nb; ++j)
            if (b[j]) ones.push_back(j);
        for (int64_t k = 0; k < n; ++k) {
            int32_t c = 0;
            for (int64_t j : ones) {
                const int64_t i = (j + k) % n;
                c += (i < na && a[i]) ? 1 : 0;
            }
            out[k] = c;
        }
        return PCS_OK;
    }
    const int64_t W = n >> 6, Wb = (nb + 63) >> 6;
    std::vector<uint64_t> A((size_t)W, 0), B((size_t)W, 0);
    for (int64_t i = 0; i < na; ++i)
        if (a[i]) A[(size_t)(i >> 6)] |= 1ull << (i & 63);
    for (int64_t i = 0; i < nb; ++i)
        if (b[i]) B[(size_t)(i >> 6)] |= 1ull << (i & 63);
    auto work = [&](int s0, int s1) {
        std::vector<uint64_t> R((size_t)W);
        for (int s = s0; s < s1; ++s) {
            // R bit i = a[(i + s) mod n]
            for (int64_t w = 0; w < W; ++w)
                R[(size_t)w] = s ? (A[(size_t)w] >> s) | (A[(size_t)((w + 1) % W)] << (64 - s)) : A[(size_t)w];
            for (int64_t q = 0; q < W; ++q) {
                int32_t c = 0;
                int64_t w2 = q;
                for (int64_t w = 0; w < Wb; ++w) {
                    c += __builtin_popcountll(R[(size_t)w2] & B[(size_t)w]);
                    if (++w2 == W) w2 = 0;
                }
                out[64 * q + s] = c;
            }
        }
    };
    unsigned nt = std::min<unsigned>(8, std::max<unsigned>(1, std::thread::hardware_concurrency()));
    if (W * Wb < (1 << 14)) nt = 1;                        // small problems: not worth a thread
    if (nt <= 1) {
        work(0, 64);
    } else {
        std::vector<std::thread> pool;
        const int per = (64 + (int)nt - 1) / (int)nt;
        for (int s0 = 0; s0 < 64; s0 += per) pool.emplace_back(work, s0, std::min(64, s0 + per));
        for (auto& t : pool) t.join();
    }
    return PCS_OK;
}

// The k largest entries by repeated arg-max (first index wins ties), as softCombiner.py:708-715 does with
// `idx[i] = argmax(x); val[i] = x[idx[i]]; x[idx[i]] = 0`.
int pcs_topk_i32(const int32_t* v, int64_t n, int32_t k, int64_t* idx_out, int32_t* val_out) {
    if (!v || !idx_out || !val_out || n < 1 || k < 1) return pcs_fail_msg(PCS_ERR_INVALID, "pcs_topk_i32: bad argument");
    std::vector<int32_t> x(v, v + n);
    for (int32_t i = 0; i < k; ++i) {
        int64_t best = 0;
        for (int64_t j = 1; j < n; ++j)
            if (x[(size_t)j] > x[(size_t)best]) best = j;
        idx_out[i] = best;
        val_out[i] = x[(size_t)best];
        x[(size_t)best] = 0;
    }
    return PCS_OK;
}

}  // extern "C"
