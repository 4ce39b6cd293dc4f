// Device kernels of the demodulator hot path (sm_100a).  See DESIGN.md for the data layout and
// the per-kernel roofline; reference line numbers refer to /root/reference/pyCuSDR/demodulator/
// cuda_kernels.cu ("kern") and demodulator_base.py ("dem_base").
#pragma once
#include <math_constants.h>
#include "fft_core.cuh"

namespace pcs {

// ---------------------------------------------------------------------------------------------
// Result block written by the device and copied to pinned host memory once per chunk.
// ---------------------------------------------------------------------------------------------
#define PCS_WINDOW_MAX 8192
#define PCS_NUM_STAGES 7
enum { PCS_STAGE_SPECTRUM = 0, PCS_STAGE_SEARCH = 1, PCS_STAGE_ESTIMATE = 2, PCS_STAGE_DEMOD_SURFACE = 3,
       PCS_STAGE_TIMING_SYMBOLS = 4, PCS_STAGE_REDUCE = 5, PCS_STAGE_BLOCK_SPECTRA = 6 };
struct DevResult {
    float best_idx;      // findDopplerEst res[0]  (kern:562,590)
    float metric_db;     // findDopplerEst res[1]  (kern:565,592)
    int low_idx;         // int(best)              (dem_base:610)
    int high_idx;        // ceil(best)             (dem_base:611)
    int shift;           // dopplerIdxlast         (dem_base:618)
    int status;          // 0 ok, 1 = NaN estimate (dem_base:625-630)
    float timing[3];     // findCodeRateAndPhase out[0..2] (kern:306-310)
    int n_sym;           // int(Nfft / spSym)      (dem_base:999)
    double sp_sym;       // dem_base:735
    double code_offset;  // dem_base:745-747
    float peak_val;      // north-star peak: max |y|^2 over (bin, mask, offset)
    int peak_bin, peak_mask, peak_offset;
    int sig_start, sig_len, noise_start, noise_len;   // circular windows of X gathered for computeSNR
    int demod_shift;     // shift actually used by the demod stage
    int xchg_timeout;    // 1 = a peer's rows never arrived (bin sharding over peer memory)
};

// ---------------------------------------------------------------------------------------------
// a6 + a7 + a8 fused (overlap-save form): for one (Doppler bin d, block b) a group
//   1. loads B chunk samples, rotates them by exp(-2 pi i s_d n / N)  (== spectrum shift by s_d,
//      kern:370) and runs a B-point forward FFT into shared memory,
//   2. for every mask m multiplies by the B-point conjugate filter spectrum (== Mk[m, k*N/B],
//      pre-scaled by N/B) and runs the inverse FFT,
//   3. reduces sum |y|^2 (kern:442-443, the 2^-18 scale is applied by the reduce kernel) and the
//      running max / arg-max over the block's valid outputs straight from registers.
// Only 12 bytes per (d, m, block) leave the SM.
// ---------------------------------------------------------------------------------------------
struct OsSearchParams {
    const float2* __restrict__ x;       // [N] chunk in HBM
    const float2* __restrict__ gb;      // [M][B] filter spectra, scaled by N/B
    const float2* __restrict__ tw;      // [B] exp(-2 pi i t / B)
    const int* __restrict__ shifts;     // [D]
    float* __restrict__ psum;           // [D][M][nblk]
    float* __restrict__ pmax;           // [D][M][nblk]
    const int* __restrict__ wblk;       // LOCATE: [D][M] block that holds the largest |y|^2 (from search_reduce_kernel)
    int* __restrict__ peak_off;         // LOCATE: [D][M] sample offset of that maximum
    const float2* __restrict__ twp;     // pass twiddle tables (PassTw<LOGB>)
    int N, D, M, nblk, V, Lpos;
    float invN;
    // shifted-filter form (FS = true): block spectra of the unrotated chunk and per-bin filter spectra
    const float2* __restrict__ xbs;     // [nblk][B]
    const float2* __restrict__ gs;      // [D][M][B] = Mk[m][(k N/B - s_d) % N] * N/B
};

struct PeakAcc {
    float sum, best;
    int idx;
    PCS_DEVINL void init() {
        sum = 0.f;
        best = -1.f;
        idx = 0x7fffffff;
    }
    PCS_DEVINL void take(float mag, int n) {
        sum += mag;
        if (mag > best || (mag == best && n < idx)) {
            best = mag;
            idx = n;
        }
    }
    PCS_DEVINL void merge(float ob, int oi) {
        if (ob > best || (ob == best && oi < idx)) {
            best = ob;
            idx = oi;
        }
    }
    PCS_DEVINL void warp_reduce() {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            float ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            merge(ob, oi);
        }
    }
};

// Doppler rotation fused into pass 0 of the forward transform: v[r] *= base * step^r.
struct RotatePre {
    float2 step;
    uint32_t shift, n_first, nmask;
    float invN;
    PCS_DEVINL void operator()(float2* v, int j) const {
        const uint32_t n = (n_first + (uint32_t)j) & nmask;
        const float2 base = unit_phasor_neg((shift * n) & nmask, invN);
        apply_twiddle_powers<16>(v, step);
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = cmul(v[r], base);
    }
};

// FS = true: the shifted-filter form (see search_fs256_kernel): no rotation and no forward transform per item; the
// block spectrum comes from block_spectra_kernel's table and the filter spectra are the bin's own.  Items are then
// ordered block-fastest so that the groups of a CTA share the bin's filter spectra in L1.
// LOCATE = false: one item = (bin, block), all masks; the epilogue keeps sum |y|^2 and max |y|^2 of the block's valid outputs
// only (a thread's 16 outputs of the last pass are t + T u, u = 0..15, so validity is a u-range per thread: no index
// arithmetic per output).  LOCATE = true: one item = (bin, mask): the block search_reduce_kernel found to hold that
// column's maximum is recomputed once to get the sample offset of the peak (lowest index wins ties).
#ifndef PCS_OS_MINB
#define PCS_OS_MINB 2
#endif
#ifndef PCS_OS_TAB
#define PCS_OS_TAB 0      // pass twiddles from tables (fft_core.cuh: PassTw) measured SLOWER on C1: 0.143-0.150 ms against 0.108 ms with
#endif                 // powers generated in registers (the 8 + 8 128-bit table loads per transform cost more registers than the chain)
template <int LOGB, int G, bool FS, bool LOCATE>
__global__ void __launch_bounds__(G * FftShape<LOGB>::T, (G * FftShape<LOGB>::T <= 256) ? PCS_OS_MINB : 1) search_os_kernel(OsSearchParams p) {
    using S = FftShape<LOGB>;
    constexpr int B = S::B, T = S::T, NW = (T + 31) / 32, NBUF = 3, R = S::RLAST, NB = 16 / R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* smem = reinterpret_cast<float2*>(smem_raw);
    const int g = threadIdx.x / T, t = threadIdx.x % T;
    float2* work0 = smem + (size_t)g * NBUF * S::WORK;
    float2* work1 = work0 + S::WORK;
    float2* xb = work1 + S::WORK;       // the item's block spectrum
    // per-group reduction scratch behind the FFT buffers: [G][M][NW] x {sum, max, idx}
    float* red = reinterpret_cast<float*>(smem + (size_t)G * NBUF * S::WORK) + (size_t)g * p.M * NW * 3;
    const int bar_id = 1 + g;

    const long long item = (long long)blockIdx.x * G + g;
    if (item >= (LOCATE ? (long long)p.D * p.M : (long long)p.nblk * p.D)) return;      // whole group leaves (own barrier id)
    int blk, d, m_lo, m_hi;
    if (LOCATE) {
        d = (int)(item / p.M);
        m_lo = (int)(item % p.M);
        m_hi = m_lo + 1;
        blk = min(max(p.wblk[item], 0), p.nblk - 1);
    } else {
        blk = FS ? (int)(item % p.nblk) : (int)(item / p.D);
        d = FS ? (int)(item / p.nblk) : (int)(item % p.D);
        m_lo = 0;
        m_hi = p.M;
    }
    const uint32_t nmask = (uint32_t)p.N - 1u;
    const int n0 = blk * p.V;

    // ---- forward transform of the rotated block -> xb (natural order) ----
    if (!FS) {
        const uint32_t shift = (uint32_t)p.shifts[d];
        const uint32_t n_first = (uint32_t)(n0 - p.Lpos) & nmask;
        RotatePre pre;
        pre.shift = shift;
        pre.n_first = n_first;
        pre.nmask = nmask;
        pre.invN = p.invN;
        pre.step = unit_phasor_neg((shift * (uint32_t)(B / 16)) & nmask, p.invN);
        auto src = [&](int i) { return __ldg(&p.x[(n_first + (uint32_t)i) & nmask]); };
        auto sink = [&](int i, float2 v, int) { xb[padi(i)] = v; };
        group_fft<LOGB, -1>(work0, work1, p.tw, t, bar_id, src, sink, pre);
    } else {
        const float2* __restrict__ xs = p.xbs + (size_t)blk * B;
#pragma unroll
        for (int r = 0; r < 16; ++r) xb[padi(t + r * T)] = __ldg(&xs[t + r * T]);
    }
    group_sync<T>(bar_id);

    const int vlen = min(p.V, p.N - n0);
    const int lane = t & 31, warp = t >> 5;
    // output index of slot (b, s) of the last pass: t + T * u with u = dft_q<R>(s) * NB + b; valid iff Lpos <= index < Lpos + vlen
    const int ulo = max(0, (p.Lpos - t + T - 1) / T), uhi = min(16, max(0, (p.Lpos + vlen - t + T - 1) / T));
    const unsigned vm = ulo < uhi ? ((uhi >= 32 ? 0xffffffffu : (1u << uhi) - 1u) & ~((1u << ulo) - 1u)) : 0u;
    for (int m = m_lo; m < m_hi; ++m) {
        const float2* __restrict__ gm = FS ? p.gs + ((size_t)d * p.M + m) * B : p.gb + (size_t)m * B;
        auto src = [&](int i) { return cmul(xb[padi(i)], __ldg(&gm[i])); };
        if (LOCATE) {
            PeakAcc acc;
            acc.init();
            auto sink = [&](int i, float2 v, int) {
                const int rel = i - p.Lpos;
                if (rel >= 0 && rel < vlen) acc.take(cabs2(v), n0 + rel);
            };
            group_fft<LOGB, +1, PCS_OS_TAB != 0>(work0, work1, PCS_OS_TAB ? p.twp : p.tw, t, bar_id, src, sink);
            acc.warp_reduce();
            if (lane == 0) {
                float* r = red + (size_t)warp * 3;
                r[1] = acc.best;
                r[2] = __int_as_float(acc.idx);
            }
        } else {
            float mg[16];
            auto sink = [&](int, float2 v, int slot) { mg[slot] = cabs2(v); };
            group_fft<LOGB, +1, PCS_OS_TAB != 0>(work0, work1, PCS_OS_TAB ? p.twp : p.tw, t, bar_id, src, sink);
            float sum = 0.f, best = 0.f;
#pragma unroll
            for (int slot = 0; slot < 16; ++slot) {
                const int u = dft_q<R>(slot % R) * NB + slot / R;
                const float mv = (vm >> u) & 1u ? mg[slot] : 0.f;
                sum += mv;
                best = fmaxf(best, mv);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sum += __shfl_xor_sync(0xffffffffu, sum, o);
                best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
            }
            if (lane == 0) {
                float* r = red + ((size_t)m * NW + warp) * 3;
                r[0] = sum;
                r[1] = best;
            }
        }
    }
    group_sync<T>(bar_id);
    if (LOCATE) {
        if (t == 0) {
            PeakAcc acc;
            acc.init();
#pragma unroll
            for (int w = 0; w < NW; ++w) acc.merge(red[(size_t)w * 3 + 1], __float_as_int(red[(size_t)w * 3 + 2]));
            p.peak_off[item] = acc.idx == 0x7fffffff ? 0 : acc.idx;
        }
    } else {
        for (int m = t; m < p.M; m += T) {
            float sum = 0.f, best = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const float* r = red + ((size_t)m * NW + w) * 3;
                sum += r[0];
                best = fmaxf(best, r[1]);
            }
            const size_t o = ((size_t)d * p.M + m) * p.nblk + blk;
            p.psum[o] = sum;
            p.pmax[o] = best;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// a6 + a7 + a8 for a FACTORISED filter bank (bank_factor.cu; the reference's FSK-2 bank, CC11xx: M = 8 filters of 3 x 128
// taps are combinations of R = 2 one-symbol basis filters).  One group per (bin, block) item:
//   1. the item's block spectrum (block_spectra_kernel's table) goes to shared memory,
//   2. R inverse transforms of (block spectrum x the bin's basis spectrum r) leave u_r = b_r * x in shared memory in natural
//      order (u_0 .. u_{R-2} in their own buffers, the last one over the block spectrum, which is dead by then),
//   3. y_m[i] = sum_j c[d][m][j] u_{sel[m][j]}[i + j S] for the block's valid outputs, two adjacent outputs per thread and
//      step (one 128-bit shared-memory load per term), sum and max of |y|^2 per mask as in search_os_kernel.
// R transforms instead of M per item; the combination costs J 128-bit loads + 2 J packed FMAs per output pair and mask.
// The offset of the maximum is recovered by search_os_kernel<.., LOCATE = true> with the unfactorised spectra.
// ---------------------------------------------------------------------------------------------
struct FbSearchParams {
    const float2* __restrict__ xbs;     // [nblk][B] block spectra of the unrotated chunk
    const float2* __restrict__ gbasis;  // [D][R][B] per-bin basis spectra
    const float2* __restrict__ coef;    // [D][M][J]  (complete binary bank: filters in code order)
    const int* __restrict__ sel;        // [M][J] basis index of every segment  (complete binary bank: [M] code -> mask)
    const float2* __restrict__ tw;      // [B] exp(-2 pi i t / B)
    float* __restrict__ psum;           // [D][M][nblk]
    float* __restrict__ pmax;           // [D][M][nblk]
    int N, D, M, nblk, V, Lpos, R, S;
};

PCS_DEVINL float2 cfma(float2 c, float2 a, float2 y) {      // y + c a: two FFMA2
    const float2 t = __ffma2_rn(make_float2(c.x, c.x), a, y);
    return __ffma2_rn(make_float2(c.y, c.y), make_float2(-a.y, a.x), t);
}

// CB = "complete binary bank": R = 2 basis filters and M = 2^J filters whose selectors enumerate all 2^J sequences (what an
// FSK-2 bank is).  The host then orders the filters by code = sum_j sel[m][j] 2^j (coef in code order, sel = code -> mask),
// every selection is static, and the combination runs from registers: per pair of outputs the 2 J values u_r[i + j S] are
// loaded ONCE (2 J 128-bit loads instead of M J) and feed 2 M independent multiply-add chains.  (The general form below
// re-reads shared memory per (mask, segment): ncu on C1 showed its epilogue at 67 % of the kernel, 1.65 IPC, bound by the
// latency of its short dependent chains.)
template <int LOGB, int G, int J, int CBM>
__global__ void __launch_bounds__(G * FftShape<LOGB>::T, (G * FftShape<LOGB>::T <= 128) ? (CBM ? 4 : 3) : (CBM && G * FftShape<LOGB>::T <= 256) ? 2 : 1)
    search_fb_kernel(FbSearchParams p) {
    constexpr bool CB = CBM != 0, TREE = CBM == 2;
    using S = FftShape<LOGB>;
    constexpr int B = S::B, T = S::T, NW = (T + 31) / 32;
    static_assert(S::NPASS == 3, "search_fb_kernel: three-pass transforms (B = 2^9 .. 2^12)");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* smem = reinterpret_cast<float2*>(smem_raw);
    const int g = threadIdx.x / T, t = threadIdx.x % T;
    // CB: three buffers in rotation (below); else work0 | work1 | xb | u_0 .. u_{R-2}
    const size_t gstride = (size_t)3 * S::WORK + (CB ? 0 : (size_t)(p.R - 1) * B);
    float2* work0 = smem + (size_t)g * gstride;
    float2* work1 = work0 + S::WORK;
    float2* xb = work1 + S::WORK;       // the item's block spectrum (padded layout)
    float2* ub = xb + S::WORK;          // (not CB) u_0 .. u_{R-2}, B each; u_{R-1} goes over xb (natural, unpadded)
    float* red = reinterpret_cast<float*>(smem + (size_t)G * gstride) + (size_t)g * p.M * NW * 2;
    float2* cs = reinterpret_cast<float2*>(reinterpret_cast<float*>(smem + (size_t)G * gstride) + (size_t)G * p.M * NW * 2) +
                 (size_t)g * p.M * J;   // the bin's M x J coefficients
    const int bar_id = 1 + g;

    const long long item = (long long)blockIdx.x * G + g;
    if (item >= (long long)p.nblk * p.D) return;      // whole group leaves (own barrier id)
    const int blk = (int)(item % p.nblk), d = (int)(item / p.nblk);
    const int n0 = blk * p.V;
    const float2* __restrict__ xs = p.xbs + (size_t)blk * B;
    for (int q = t; q < p.M * J; q += T) cs[q] = __ldg(&p.coef[(size_t)d * p.M * J + q]);
    if constexpr (!CB) {
#pragma unroll
        for (int r = 0; r < 16; ++r) xb[padi(t + r * T)] = __ldg(&xs[t + r * T]);
        group_sync<T>(bar_id);
    }
    const float2 *u0 = ub, *u1 = xb;    // CB: where u_0 and u_1 end up
    if constexpr (CB) {
        // R = 2 transforms through THREE buffers (52 KB per 2048-point group: four 128-thread CTAs per SM instead of three):
        //   u_0:  xb x G_0 -> work0 -> work1 -> work0 (natural order);   u_1:  xb x G_1 -> work1 -> xb -> work1 (natural order)
        // Every pass reads one buffer and writes another one that no thread still reads: the group barrier after each pass
        // (one more than back-to-back group_fft calls need: the one after u_0's last pass frees work1) orders them.
        const float2* __restrict__ g0 = p.gbasis + (size_t)d * 2 * B;
        const float2* __restrict__ g1 = g0 + B;
        auto ld0 = [&](int i) { return work0[padi(i)]; };
        auto ld1 = [&](int i) { return work1[padi(i)]; };
        auto ldx = [&](int i) { return xb[padi(i)]; };
        auto st0 = [&](int i, float2 v, int) { work0[padi(i)] = v; };
        auto st1 = [&](int i, float2 v, int) { work1[padi(i)] = v; };
        auto stx = [&](int i, float2 v, int) { xb[padi(i)] = v; };
        auto nat0 = [&](int i, float2 v, int) { work0[i] = v; };
        auto nat1 = [&](int i, float2 v, int) { work1[i] = v; };
        // pass 0 reads inputs t + r T: exactly the elements of the block spectrum this thread would stage, so u_0's pass 0
        // takes them straight from global memory and parks them in xb for u_1's pass 0 (same thread: no barrier in between)
        auto src0 = [&](int i) {
            const float2 xv = __ldg(&xs[i]);
            xb[padi(i)] = xv;
            return cmul(xv, __ldg(&g0[i]));
        };
        auto src1 = [&](int i) { return cmul(xb[padi(i)], __ldg(&g1[i])); };
        fft_pass<LOGB, 16, 0, +1>(t, p.tw, src0, st0);
        group_sync<T>(bar_id);
        fft_pass<LOGB, 16, 4, +1>(t, p.tw, ld0, st1);
        group_sync<T>(bar_id);
        fft_pass<LOGB, S::RLAST, 8, +1>(t, p.tw, ld1, nat0);
        group_sync<T>(bar_id);
        fft_pass<LOGB, 16, 0, +1>(t, p.tw, src1, st1);
        group_sync<T>(bar_id);
        fft_pass<LOGB, 16, 4, +1>(t, p.tw, ld1, stx);
        group_sync<T>(bar_id);
        fft_pass<LOGB, S::RLAST, 8, +1>(t, p.tw, ldx, nat1);
        u0 = work0;
        u1 = work1;
    } else {
        for (int r = 0; r < p.R; ++r) {
            const float2* __restrict__ gm = p.gbasis + ((size_t)d * p.R + r) * B;
            float2* __restrict__ dst = r == p.R - 1 ? xb : ub + (size_t)r * B;
            auto src = [&](int i) { return cmul(xb[padi(i)], __ldg(&gm[i])); };
            auto sink = [&](int i, float2 v, int) { dst[i] = v; };
            group_fft<LOGB, +1>(work0, work1, p.tw, t, bar_id, src, sink);
        }
    }
    group_sync<T>(bar_id);

    const int lo = p.Lpos, hi = p.Lpos + min(p.V, p.N - n0);     // valid outputs of the block: lo <= i < hi
    const int lane = t & 31, warp = t >> 5;
    if constexpr (CB) {
        constexpr int MC = 1 << J;
        // TREE: c[code][j] depends on the low j + 1 bits of the code only (FSK-2: the phase a segment starts with is set by
        // the symbols before it), so the partial sums over segments 0..j are shared by the codes with the same prefix:
        // 2 + 4 + .. + 2^J multiply-adds per output instead of J 2^J, and that many coefficients in registers.
        float2 c[MC][J];
#pragma unroll
        for (int code = 0; code < MC; ++code)
#pragma unroll
            for (int j = 0; j < J; ++j)
                if (!TREE || code < (2 << j)) c[code][j] = cs[code * J + j];
        float sum[MC], best[MC];
#pragma unroll
        for (int code = 0; code < MC; ++code) sum[code] = best[code] = 0.f;
        const int nq = (hi + 2 * T - 1) / (2 * T);      // steps that hold a valid output (uniform over the group)
        const int last = (hi - 1) & ~1;                 // last pair with a valid output: loads beyond it are clamped to it
#pragma unroll
        for (int q = 0; q < B / (2 * T); ++q) {
            if (q < nq) {
                const int i0 = 2 * (t + T * q), ic = min(i0, last);
                float4 P[J][2];
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    P[j][0] = *reinterpret_cast<const float4*>(u0 + ic + j * p.S);
                    P[j][1] = *reinterpret_cast<const float4*>(u1 + ic + j * p.S);
                }
                const bool v0 = i0 >= lo && i0 < hi, v1 = i0 + 1 >= lo && i0 + 1 < hi;
                float2 y0[MC], y1[MC];
                if constexpr (TREE) {
#pragma unroll
                    for (int j = 0; j < J; ++j)
#pragma unroll
                        for (int pre = (2 << j) - 1; pre >= 0; --pre) {      // downwards: y[pre] of level j - 1 is read before it is replaced
                            const float4 a = P[j][(pre >> j) & 1];
                            if (j == 0) {
                                y0[pre] = cmul(c[pre][0], make_float2(a.x, a.y));
                                y1[pre] = cmul(c[pre][0], make_float2(a.z, a.w));
                            } else {
                                y0[pre] = cfma(c[pre][j], make_float2(a.x, a.y), y0[pre & ((1 << j) - 1)]);
                                y1[pre] = cfma(c[pre][j], make_float2(a.z, a.w), y1[pre & ((1 << j) - 1)]);
                            }
                        }
                } else {
#pragma unroll
                    for (int code = 0; code < MC; ++code) {
                        float4 a = P[0][code & 1];
                        y0[code] = cmul(c[code][0], make_float2(a.x, a.y));
                        y1[code] = cmul(c[code][0], make_float2(a.z, a.w));
#pragma unroll
                        for (int j = 1; j < J; ++j) {
                            a = P[j][(code >> j) & 1];
                            y0[code] = cfma(c[code][j], make_float2(a.x, a.y), y0[code]);
                            y1[code] = cfma(c[code][j], make_float2(a.z, a.w), y1[code]);
                        }
                    }
                }
#pragma unroll
                for (int code = 0; code < MC; ++code) {
                    const float m0 = v0 ? cabs2(y0[code]) : 0.f, m1 = v1 ? cabs2(y1[code]) : 0.f;
                    sum[code] += m0;
                    sum[code] += m1;
                    best[code] = fmaxf(best[code], fmaxf(m0, m1));
                }
            }
        }
#pragma unroll
        for (int code = 0; code < MC; ++code) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sum[code] += __shfl_xor_sync(0xffffffffu, sum[code], o);
                best[code] = fmaxf(best[code], __shfl_xor_sync(0xffffffffu, best[code], o));
            }
            if (lane == 0) {
                float* r = red + ((size_t)code * NW + warp) * 2;
                r[0] = sum[code];
                r[1] = best[code];
            }
        }
    } else {
        for (int m = 0; m < p.M; ++m) {
            float2 c[J];
            const float2* up[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                c[j] = cs[m * J + j];
                const int r = __ldg(&p.sel[m * J + j]);
                up[j] = (r == p.R - 1 ? xb : ub + (size_t)r * B) + j * p.S;
            }
            float sum = 0.f, best = 0.f;
#pragma unroll
            for (int q = 0; q < B / (2 * T); ++q) {
                const int i0 = 2 * (t + T * q);
                if (i0 < hi) {
                    float4 a = *reinterpret_cast<const float4*>(up[0] + i0);
                    float2 y0 = cmul(c[0], make_float2(a.x, a.y)), y1 = cmul(c[0], make_float2(a.z, a.w));
#pragma unroll
                    for (int j = 1; j < J; ++j) {
                        a = *reinterpret_cast<const float4*>(up[j] + i0);
                        y0 = cfma(c[j], make_float2(a.x, a.y), y0);
                        y1 = cfma(c[j], make_float2(a.z, a.w), y1);
                    }
                    const float m0 = i0 >= lo ? cabs2(y0) : 0.f;
                    const float m1 = (i0 + 1 >= lo && i0 + 1 < hi) ? cabs2(y1) : 0.f;
                    sum += m0;
                    sum += m1;
                    best = fmaxf(best, fmaxf(m0, m1));
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sum += __shfl_xor_sync(0xffffffffu, sum, o);
                best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
            }
            if (lane == 0) {
                float* r = red + ((size_t)m * NW + warp) * 2;
                r[0] = sum;
                r[1] = best;
            }
        }
    }
    group_sync<T>(bar_id);
    for (int m = t; m < p.M; m += T) {        // CB: m is a code, sel[code] its mask
        float sum = 0.f, best = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float* r = red + ((size_t)m * NW + w) * 2;
            sum += r[0];
            best = fmaxf(best, r[1]);
        }
        const int mask = CB ? __ldg(&p.sel[m]) : m;
        const size_t o = ((size_t)d * p.M + mask) * p.nblk + blk;
        p.psum[o] = sum;
        p.pmax[o] = best;
    }
}

// Block spectra of the unrotated chunk for the shifted-filter form of the generic kernel: one group per block.
template <int LOGB, int G>
__global__ void __launch_bounds__(G * FftShape<LOGB>::T) block_spectra_kernel(const float2* __restrict__ x,
                                                                               const float2* __restrict__ tw,
                                                                               float2* __restrict__ xbs, int N, int nblk, int V,
                                                                               int Lpos) {
    using S = FftShape<LOGB>;
    constexpr int B = S::B, T = S::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* smem = reinterpret_cast<float2*>(smem_raw);
    const int g = threadIdx.x / T, t = threadIdx.x % T;
    float2* work0 = smem + (size_t)g * 2 * S::WORK;
    float2* work1 = work0 + S::WORK;
    const int blk = blockIdx.x * G + g;
    if (blk >= nblk) return;
    const uint32_t nmask = (uint32_t)N - 1u, n_first = (uint32_t)(blk * V - Lpos) & nmask;
    float2* __restrict__ out = xbs + (size_t)blk * B;
    auto src = [&](int i) { return __ldg(&x[(n_first + (uint32_t)i) & nmask]); };
    auto sink = [&](int i, float2 v, int) { out[i] = v; };
    group_fft<LOGB, -1>(work0, work1, tw, t, 1 + g, src, sink);
}

// G[d][m][k] = Mk[m][(k N/B - s_d) % N] * N/B, natural order (once per handle).
__global__ void shifted_filters_kernel(const float2* __restrict__ masks, const int* __restrict__ shifts,
                                       float2* __restrict__ gs, int N, int logB, int D, int M) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ((long long)D * M) << logB) return;
    const int k = (int)(idx & ((1 << logB) - 1));
    const int dm = (int)(idx >> logB), m = dm % M, d = dm / M;
    const uint32_t nmask = (uint32_t)N - 1u, dec = (uint32_t)(N >> logB);
    const float scale = (float)dec;
    const float2 a = masks[(size_t)m * N + (((uint32_t)k * dec - (uint32_t)shifts[d]) & nmask)];
    gs[idx] = make_float2(a.x * scale, a.y * scale);
}

// ---------------------------------------------------------------------------------------------
// a6 + a7 + a8 fused, register-resident form for short filters (B = 256 = 16 x 16).
// One group = 16 lanes (half a warp), 16 points per lane.  With a two-pass radix-16 autosort transform lane t
// consumes inputs t + 16 r and produces outputs t + 16 q, so the block spectrum a forward transform leaves in a
// lane's registers is exactly what pass 0 of the inverse transforms wants from that lane: the spectrum never
// touches shared memory, every transform needs ONE 2 KB shared-memory exchange, the 15 inter-pass twiddles
// W_256^(t r) are per-lane constants held in registers, and the filter spectra arrive as float4 (layout
// [m][r/2][t][r%2]).  The epilogue keeps sum |y|^2 and max |y|^2 only; the offset of the maximum is recovered
// by peak_locate256_kernel for the one winning block of each (bin, mask).
// Partials are stored [block][bin][mask] so a group's M results are one contiguous store.
// ---------------------------------------------------------------------------------------------
struct Os256Params {
    const float2* __restrict__ x;       // [N] chunk in HBM
    const float4* __restrict__ gperm;   // [M][8][16] float4 = filter spectra (x N/256), lane-major pairs
    const float2* __restrict__ tw;      // [256] exp(-2 pi i t / 256)
    const int* __restrict__ shifts;     // [D]
    float* __restrict__ psum;           // [nblk][D][M]
    float* __restrict__ pmax;           // [nblk][D][M]
    int N, D, M, nblk, V, Lpos;
    float invN;
};

PCS_DEVINL float2 cmulc(float2 a, float2 b) {   // a * conj(b) = fma2(a, b.x, (a.y b.y, -(a.x b.y))): FMUL2 + FFMA2
    const float2 t = __fmul2_rn(make_float2(b.y, b.y), make_float2(a.y, a.x));
    return __ffma2_rn(a, make_float2(b.x, b.x), make_float2(t.x, -t.y));
}

// 256-point transform of one 16-lane group. In: v[r] = in[t + 16 r]. Out: slot s holds out[t + 16 dft_q<16>(s)].
// tw[r] = exp(-2 pi i t r / 256).  buf: the group's 272-float2 exchange buffer.
template <int DIR>
PCS_DEVINL void fft256_regs(float2* v, float2* buf, const float2* tw, int t) {
    Dft<16, DIR>::run(v);
    __syncwarp();                      // previous transform's loads are done before the buffer is overwritten
    // Exchange through the group's 272-float2 buffer: element (storing lane r, output q) = natural index 16 r + q lives
    // at 34 (r >> 1) + 2 q + (r & 1), so that the loading lane t finds its inputs t + 16 (2j) and t + 16 (2j + 1) side by
    // side: 16 STS.64 + 8 LDS.128 per transform, both conflict-free (row stride 68 words = 4 mod 32).
    {
        float2* row = buf + 34 * (t >> 1) + (t & 1);
#pragma unroll
        for (int s = 0; s < 16; ++s) row[2 * dft_q<16>(s)] = v[s];
    }
    __syncwarp();
    {
        const float4* col = reinterpret_cast<const float4*>(buf) + t;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 q = col[17 * j];
            v[2 * j] = make_float2(q.x, q.y);
            v[2 * j + 1] = make_float2(q.z, q.w);
        }
    }
#pragma unroll
    for (int r = 1; r < 16; ++r) v[r] = DIR < 0 ? cmul(v[r], tw[r]) : cmulc(v[r], tw[r]);
    Dft<16, DIR>::run(v);
}

struct Os256Item {
    int blk, d, n0, vlen;
    uint32_t shift, n_first, nmask;
};

PCS_DEVINL Os256Item os256_item(const Os256Params& p, long long item) {
    Os256Item it;
    it.blk = (int)(item / p.D);
    it.d = (int)(item % p.D);
    it.nmask = (uint32_t)p.N - 1u;
    it.shift = (uint32_t)p.shifts[it.d];
    it.n0 = it.blk * p.V;
    it.n_first = (uint32_t)(it.n0 - p.Lpos) & it.nmask;
    it.vlen = min(p.V, p.N - it.n0);
    return it;
}

// Block spectrum of the Doppler-rotated block (== spectrum shift by s_d, kern:370), natural order r -> X[t + 16 r].
PCS_DEVINL void os256_block_spectrum(const Os256Params& p, const Os256Item& it, float2* buf, const float2* tw, int t,
                                     float2* xb) {
    float2 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = __ldg(&p.x[(it.n_first + (uint32_t)(t + 16 * r)) & it.nmask]);
    const uint32_t n = (it.n_first + (uint32_t)t) & it.nmask;
    const float2 base = unit_phasor_neg((it.shift * n) & it.nmask, p.invN);
    const float2 step = unit_phasor_neg((it.shift * 16u) & it.nmask, p.invN);
    apply_twiddle_powers<16>(v, step);
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = cmul(v[r], base);
    fft256_regs<-1>(v, buf, tw, t);
#pragma unroll
    for (int r = 0; r < 16; ++r) xb[r] = v[dft_q<16>(r)];     // dft_q<16> is an involution
}

// y = IFFT(xb * G_m) for one filter, left in slot order (slot s <-> output index t + 16 dft_q<16>(s)).
PCS_DEVINL void os256_filter(const Os256Params& p, int m, const float2* xb, float2* buf, const float2* tw, int t,
                             float2* v) {
    const float4* __restrict__ g4 = p.gperm + (size_t)m * 128 + t;
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
        const float4 g = __ldg(&g4[rr * 16]);
        v[2 * rr] = cmul(xb[2 * rr], make_float2(g.x, g.y));
        v[2 * rr + 1] = cmul(xb[2 * rr + 1], make_float2(g.z, g.w));
    }
    fft256_regs<+1>(v, buf, tw, t);
}

// Same with the block spectrum parked in shared memory ([r][16 lanes] float2) instead of 32 registers per lane.
PCS_DEVINL void os256_filter_smem(const Os256Params& p, int m, const float2* xbs, float2* buf, const float2* tw, int t,
                                  float2* v) {
    const float4* __restrict__ g4 = p.gperm + (size_t)m * 128 + t;
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
        const float4 g = __ldg(&g4[rr * 16]);
        v[2 * rr] = cmul(xbs[(2 * rr) * 16 + t], make_float2(g.x, g.y));
        v[2 * rr + 1] = cmul(xbs[(2 * rr + 1) * 16 + t], make_float2(g.z, g.w));
    }
    fft256_regs<+1>(v, buf, tw, t);
}

PCS_DEVINL unsigned os256_valid_mask(const Os256Params& p, const Os256Item& it, int t) {
    unsigned vm = 0;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const int rel = t + 16 * dft_q<16>(s) - p.Lpos;
        if (rel >= 0 && rel < it.vlen) vm |= 1u << s;
    }
    return vm;
}

// XBS: keep the block spectrum in shared memory (fewer registers -> more resident warps) instead of registers.
template <int G, bool XBS>     // G = groups (half warps) per CTA
__global__ void __launch_bounds__(G * 16, XBS ? 40 / G : 32 / G) search_os256_kernel(Os256Params p) {
    __shared__ __align__(16) float2 sbuf[G][272];
    __shared__ float2 sxb[XBS ? G : 1][XBS ? 256 : 1];
    extern __shared__ float s_acc[];          // [G groups][2][M][17]: per-lane sum / max of every mask
    const int t = threadIdx.x & 15, g = threadIdx.x >> 4;
    float2* buf = sbuf[g];
    float* acc_sum = s_acc + (size_t)g * 2 * p.M * 17;
    float* acc_max = acc_sum + (size_t)p.M * 17;
    float2 tw[16];
#pragma unroll
    for (int r = 1; r < 16; ++r) tw[r] = __ldg(&p.tw[(t * r) & 255]);
    tw[0] = make_float2(1.f, 0.f);
    const long long total = (long long)p.nblk * p.D;
    long long item = (long long)blockIdx.x * G + g;
    const bool live = item < total;
    if (!live) item = total - 1;          // keep the warp convergent; results of the duplicate are dropped
    const Os256Item it = os256_item(p, item);
    float2 xb[16];
    os256_block_spectrum(p, it, buf, tw, t, xb);
    if (XBS) {
#pragma unroll
        for (int r = 0; r < 16; ++r) sxb[XBS ? g : 0][(XBS ? r * 16 + t : 0)] = xb[r];
        __syncwarp();
    }
    const unsigned vm = os256_valid_mask(p, it, t);
#pragma unroll 1
    for (int m = 0; m < p.M; ++m) {
        float2 v[16];
        if (XBS) os256_filter_smem(p, m, sxb[XBS ? g : 0], buf, tw, t, v);
        else os256_filter(p, m, xb, buf, tw, t, v);
        float sum = 0.f, best = 0.f;
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            const float mag = (vm >> s) & 1u ? cabs2(v[s]) : 0.f;
            sum += mag;
            best = fmaxf(best, mag);
        }
        // no cross-lane traffic inside the loop: the per-lane partials are parked in shared memory
        acc_sum[m * 17 + t] = sum;
        acc_max[m * 17 + t] = best;
    }
    __syncwarp();
    // lane m folds the 16 lanes of mask m in lane order (fixed order -> bit-reproducible) and stores the pair
    for (int m = t; m < p.M; m += 16) {
        float sum = 0.f, best = 0.f;
#pragma unroll
        for (int l = 0; l < 16; ++l) {
            sum += acc_sum[m * 17 + l];
            best = fmaxf(best, acc_max[m * 17 + l]);
        }
        if (live) {
            const size_t o = ((size_t)it.blk * p.D + it.d) * p.M + m;
            p.psum[o] = sum;
            p.pmax[o] = best;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// a6 + a7 + a8 fused, "shifted filter" form of the 256-point search (default for M <= 16).
// The reference's spectrum shift can be charged to the filter instead of the chunk:
//   y[d,m,n] = sum_k X[(k+s_d)%N] Mk[m,k] e^{+2 pi i k n/N} = e^{-2 pi i s_d n/N} * sum_k X[k] Mk[m,(k-s_d)%N] e^{+2 pi i k n/N},
// and the unit-modulus factor drops out of |y|^2 (kern:339-373 + kern:421-480 only ever use |y|^2).  So the block
// spectra of the *unrotated* chunk are computed once per chunk (block_spectra256_kernel: nblk forward transforms instead
// of D * nblk) and every Doppler bin gets its own set of 256-point filter spectra G[d][m][k] = Mk[m][(k N/256 - s_d) % N] N/256,
// a pure gather from the protocol's own spectra done once per handle (shifted_filters256_kernel).  One item = one
// (bin, block): load 2 KB of block spectrum, then M x (filter product + inverse transform + |y|^2 sum / max).
// A CTA works on items_per_cta consecutive items of the (bin-major) item space; the bin's filter spectra sit in shared
// memory and are reloaded when the CTA crosses into the next bin.
// ---------------------------------------------------------------------------------------------
struct Fs256Params {
    const float4* __restrict__ xbs;     // [nblk][8][16] float4: block spectra, lane-major pairs (X[t+32rr], X[t+32rr+16])
    const float4* __restrict__ gs;      // [D][M][8][16] float4: Doppler-shifted filter spectra (x N/256), same layout
    const float2* __restrict__ tw;      // [256] exp(-2 pi i t / 256)
    float* __restrict__ psum;           // [D][nblk][M]  per-item partials, bin-major: a bin's partials are contiguous
    float* __restrict__ pmax;           // [D][nblk][M]
    int N, D, M, nblk, V, Lpos, items_per_cta;
    // fused finish (see fs256_finish_bin): the CTA that completes a bin reduces and locates it
    unsigned int* __restrict__ bin_count;   // [D] items of the bin finished so far in this launch (self-resetting)
    unsigned int* __restrict__ bins_done;   // bins finished in this launch (self-resetting)
    float* __restrict__ Efull;          // [D][M] rows of this launch's bin range (kern:442 scale); may be peer memory
    float* __restrict__ peak_val;       // [D][M]
    int* __restrict__ peak_off;         // [D][M]
    unsigned long long* arrival_flag;   // when non-null: raised (system-scope release) once every bin's rows are stored
    unsigned long long* ack_flag;       // optional second flag raised with it ("this rank has finished reading the chunk")
    unsigned long long arrival_value;
};

// Block spectra of the chunk: block b = FFT_256(x[(b V - Lpos + i) % N], i = 0..255), one 16-lane group per block.
template <int G>      // groups (blocks) per CTA
__global__ void __launch_bounds__(G * 16) block_spectra256_kernel(const float2* __restrict__ x, const float2* __restrict__ twg,
                                                                  float4* __restrict__ xbs, int N, int nblk, int V, int Lpos) {
    __shared__ __align__(16) float2 sbuf[G][272];
    const int t = threadIdx.x & 15, g = threadIdx.x >> 4;
    float2 tw[16];
#pragma unroll
    for (int r = 1; r < 16; ++r) tw[r] = __ldg(&twg[(t * r) & 255]);
    tw[0] = make_float2(1.f, 0.f);
    int blk = blockIdx.x * G + g;
    const bool live = blk < nblk;
    if (!live) blk = nblk - 1;
    const uint32_t nmask = (uint32_t)N - 1u, n_first = (uint32_t)(blk * V - Lpos) & nmask;
    float2 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = __ldg(&x[(n_first + (uint32_t)(t + 16 * r)) & nmask]);
    fft256_regs<-1>(v, sbuf[g], tw, t);
    if (live) {
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
            const float2 a = v[dft_q<16>(2 * rr)], b = v[dft_q<16>(2 * rr + 1)];     // natural order X[t + 16 r]
            xbs[(size_t)blk * 128 + rr * 16 + t] = make_float4(a.x, a.y, b.x, b.y);
        }
    }
}

// G[d][m][k] = Mk[m][(k N/256 - s_d) % N] * N/256 in the lane-major pair layout (once per handle).
__global__ void shifted_filters256_kernel(const float2* __restrict__ masks, const int* __restrict__ shifts,
                                          float4* __restrict__ gs, int N, int D, int M) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)D * M * 128) return;
    const int t = (int)(idx & 15), rr = (int)((idx >> 4) & 7);
    const int dm = (int)(idx >> 7), m = dm % M, d = dm / M;
    const uint32_t nmask = (uint32_t)N - 1u, dec = (uint32_t)(N >> 8), s = (uint32_t)shifts[d];
    const float scale = (float)dec;
    const float2 a = masks[(size_t)m * N + (((uint32_t)(t + 32 * rr) * dec - s) & nmask)];
    const float2 b = masks[(size_t)m * N + (((uint32_t)(t + 32 * rr + 16) * dec - s) & nmask)];
    gs[idx] = make_float4(a.x * scale, a.y * scale, b.x * scale, b.y * scale);
}

// v = IFFT_256(xb * g) with g read as float4 pairs at stride 16 from g4 (shared or global memory, already offset by t).
PCS_DEVINL void fs256_filter(const float4* g4, const float2* xb, float2* buf, const float2* tw, int t, float2* v) {
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
        const float4 g = g4[rr * 16];
        v[2 * rr] = cmul(xb[2 * rr], make_float2(g.x, g.y));
        v[2 * rr + 1] = cmul(xb[2 * rr + 1], make_float2(g.z, g.w));
    }
    fft256_regs<+1>(v, buf, tw, t);
}

PCS_DEVINL unsigned fs256_valid_mask(int Lpos, int vlen, int t) {
    unsigned vm = 0;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const int rel = t + 16 * dft_q<16>(s) - Lpos;
        if (rel >= 0 && rel < vlen) vm |= 1u << s;
    }
    return vm;
}

// Fused finish of one Doppler bin, run by the CTA whose items completed the bin (every thread of the CTA calls it):
//   1. E[d][m] = 2^-18 * sum over the bin's nblk block partials and the (value, block) of the largest |y|^2, in a fixed
//      order that depends on nblk only -- 64 interleaved lanes per column (lane j takes blocks j, j + 64, ... in
//      increasing order), then the 64 lane results in lane order -- so E is bit-identical for every CTA tiling, group count
//      and bin sharding (kern:421-480 uses float atomics: not reproducible even run to run).  The partials are L2-resident;
//      what this step costs is L2 latency, so every thread keeps 8 independent 128-bit loads in flight;
//   2. the offset of the maximum inside the winning block of every mask (lowest sample index wins ties): the block is
//      recomputed with the bin's filter spectra, which are still in shared memory;
//   3. the bin's row of the three tables is stored (into the owner's exchange region over NVLink when the search is bin-
//      sharded), and the CTA that finishes the launch's last bin raises the arrival flag with a system-scope release.
// No separate reduction / locate / flag kernels remain on the per-chunk path of the shifted-filter search.
#define PCS_FIN_LANES 64
struct FinAcc {
    float sum, best;
    int bb;
    PCS_DEVINL void init() { sum = 0.f; best = -1.f; bb = 0x7fffffff; }
    PCS_DEVINL void take(float s, float v, int b) {
        sum += s;
        if (v > best) { best = v; bb = b; }
    }
};
template <int G>
PCS_DEVINL void fs256_finish_bin(const Fs256Params& p, int d, const float4* s_g, float2* buf, const float2* tw,
                                 float (*s_red)[PCS_FIN_LANES][16], int* s_wblk) {
    constexpr int NT = G * 16;
    const int tid = threadIdx.x, t = tid & 15, g = tid >> 4;
    const int MQ = (p.M + 3) >> 2;                                   // column quads
    const bool vec = (p.M & 3) == 0;                                 // 128-bit loads need 16-byte aligned rows
    const float* __restrict__ ps = p.psum + (size_t)d * p.nblk * p.M;
    const float* __restrict__ pm = p.pmax + (size_t)d * p.nblk * p.M;
    for (int w = tid; w < PCS_FIN_LANES * MQ; w += NT) {             // work item = (lane, column quad)
        const int lane = w / MQ, c0 = (w % MQ) * 4;
        FinAcc a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k].init();
        constexpr int U = 4, STEP = PCS_FIN_LANES;
        for (int b = lane; b < p.nblk; b += U * STEP) {
            float4 s4[U], v4[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {                            // written by other SMs in this launch: bypass L1
                const int bu = b + u * STEP;
                s4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                v4[u] = make_float4(-1.f, -1.f, -1.f, -1.f);
                if (bu < p.nblk) {
                    const size_t o = (size_t)bu * p.M + c0;
                    if (vec) {
                        s4[u] = __ldcg(reinterpret_cast<const float4*>(ps + o));
                        v4[u] = __ldcg(reinterpret_cast<const float4*>(pm + o));
                    } else {
                        if (c0 + 0 < p.M) { s4[u].x = __ldcg(ps + o + 0); v4[u].x = __ldcg(pm + o + 0); }
                        if (c0 + 1 < p.M) { s4[u].y = __ldcg(ps + o + 1); v4[u].y = __ldcg(pm + o + 1); }
                        if (c0 + 2 < p.M) { s4[u].z = __ldcg(ps + o + 2); v4[u].z = __ldcg(pm + o + 2); }
                        if (c0 + 3 < p.M) { s4[u].w = __ldcg(ps + o + 3); v4[u].w = __ldcg(pm + o + 3); }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {                            // (an absent block adds + 0.0f and never wins the max)
                const int bu = b + u * STEP;
                a[0].take(s4[u].x, v4[u].x, bu);
                a[1].take(s4[u].y, v4[u].y, bu);
                a[2].take(s4[u].z, v4[u].z, bu);
                a[3].take(s4[u].w, v4[u].w, bu);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (c0 + k < p.M) {
                s_red[0][lane][c0 + k] = a[k].sum;
                s_red[1][lane][c0 + k] = a[k].best;
                s_red[2][lane][c0 + k] = __int_as_float(a[k].bb);
            }
        }
    }
    __syncthreads();
    if (tid < p.M) {
        float sum = s_red[0][0][tid], best = s_red[1][0][tid];
        int bb = __float_as_int(s_red[2][0][tid]);
#pragma unroll 8
        for (int l = 1; l < PCS_FIN_LANES; ++l) {
            sum += s_red[0][l][tid];
            const float v = s_red[1][l][tid];
            const int bl = __float_as_int(s_red[2][l][tid]);
            if (v > best || (v == best && bl < bb)) { best = v; bb = bl; }
        }
        if (bb == 0x7fffffff) bb = 0;
        p.Efull[(size_t)d * p.M + tid] = sum * (1.0f / 262144.0f);      // kern:442 (exact power-of-two scale)
        p.peak_val[(size_t)d * p.M + tid] = best;
        s_wblk[tid] = bb;
    }
    __syncthreads();                               // (also: the lane table is dead before the exchange buffers are reused)
    for (int m0 = 0; m0 < p.M; m0 += G) {          // group g recomputes the winning block of mask m0 + g
        int mm = m0 + g;
        const bool live = mm < p.M;
        if (!live) mm = p.M - 1;                   // keep both halves of every warp convergent (full-mask shuffles below)
        const int wblk = s_wblk[mm];
        float2 xb[16], v[16];
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
            const float4 q = __ldg(&p.xbs[(size_t)wblk * 128 + rr * 16 + t]);
            xb[2 * rr] = make_float2(q.x, q.y);
            xb[2 * rr + 1] = make_float2(q.z, q.w);
        }
        fs256_filter(s_g + (size_t)mm * 128 + t, xb, buf, tw, t, v);
        const int n0 = wblk * p.V, vlen = min(p.V, p.N - n0);
        float best = -1.f;
        int idx = 0x7fffffff;
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            const int rel = t + 16 * dft_q<16>(s) - p.Lpos;
            if (rel >= 0 && rel < vlen) {
                const float mag = cabs2(v[s]);
                const int n = n0 + rel;
                if (mag > best || (mag == best && n < idx)) { best = mag; idx = n; }
            }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
        }
        if (live && t == 0) p.peak_off[(size_t)d * p.M + mm] = idx == 0x7fffffff ? 0 : idx;
    }
    // publish: every thread's table stores are fenced (system scope when they went to a peer), then one thread counts the
    // bin; the CTA that counts the last one raises the flag
    if (p.arrival_flag != nullptr) __threadfence_system(); else __threadfence();
    __syncthreads();
    if (tid == 0) {
        p.bin_count[d] = 0;                        // ready for the next launch on these buffers
        const unsigned int prev = atomicAdd(p.bins_done, 1u);
        if (prev == (unsigned int)p.D - 1u) {
            *p.bins_done = 0;
            if (p.arrival_flag != nullptr) {
                __threadfence_system();
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.arrival_flag), "l"(p.arrival_value) : "memory");
                if (p.ack_flag != nullptr)
                    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.ack_flag), "l"(p.arrival_value) : "memory");
            }
        }
    }
}

// Register budget of the kernel below.  The mask loop is bound by the FMA pipe and its speed depends on the instruction
// schedule ptxas finds; measured on C2 (profiles/r02_kernel_variants.md): capped at 112 registers 0.761 ms, at the 128 that
// __launch_bounds__(128, 4) allows 0.771 ms, at 104 (spills) 0.776 ms, with the finish as a separate function 0.81 ms.
#ifndef PCS_FS_MAXREG
#define PCS_FS_MAXREG 112
#endif

template <int G, int WARPS_PER_SM = 16>     // G = groups (half warps) per CTA; WARPS_PER_SM = 20 is the 96-register tuning build
__global__ void __maxnreg__(WARPS_PER_SM > 16 ? 96 : PCS_FS_MAXREG) search_fs256_kernel(Fs256Params p) {
    // the finish's lane table (12 KB) lives in the groups' exchange buffers when they are large enough: the two are never
    // in use at the same time (every group is past its last transform when a bin is finished)
    constexpr size_t RED_BYTES = sizeof(float) * 3 * PCS_FIN_LANES * 16;
    constexpr bool ALIAS = sizeof(float2) * G * 272 >= RED_BYTES;
    __shared__ __align__(16) float2 sbuf[G][272];
    __shared__ __align__(16) float s_red_own[ALIAS ? 4 : 3 * PCS_FIN_LANES * 16];
    float (*s_red)[PCS_FIN_LANES][16] =
        reinterpret_cast<float (*)[PCS_FIN_LANES][16]>(ALIAS ? reinterpret_cast<float*>(&sbuf[0][0]) : s_red_own);
    __shared__ int s_wblk[16];
    __shared__ int s_last;
    extern __shared__ float4 s_dyn[];         // [M][128] float4 filter spectra of the current bin | [G][2][M][17] float partials
    float4* s_g = s_dyn;
    const int t = threadIdx.x & 15, g = threadIdx.x >> 4;
    float2* buf = sbuf[g];
    float* acc_sum = reinterpret_cast<float*>(s_dyn + (size_t)p.M * 128) + (size_t)g * 2 * p.M * 17;
    float* acc_max = acc_sum + (size_t)p.M * 17;
    float2 tw[16];
#pragma unroll
    for (int r = 1; r < 16; ++r) tw[r] = __ldg(&p.tw[(t * r) & 255]);
    tw[0] = make_float2(1.f, 0.f);
    const unsigned vm_full = fs256_valid_mask(p.Lpos, p.V, t);
    const long long total = (long long)p.nblk * p.D;
    long long cur = (long long)blockIdx.x * p.items_per_cta;
    const long long end = min(cur + (long long)p.items_per_cta, total);
    while (cur < end) {
        const int d = (int)(cur / p.nblk);
        const int b0 = (int)(cur - (long long)d * p.nblk);
        const int nb = (int)min((long long)(p.nblk - b0), end - cur);
        __syncthreads();                      // everyone is done with the previous bin's spectra
        {
            const float4* __restrict__ src = p.gs + (size_t)d * p.M * 128;
            for (int i = threadIdx.x; i < p.M * 128; i += G * 16) s_g[i] = __ldg(&src[i]);
        }
        __syncthreads();
        const int iters = (nb + G - 1) / G;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            int b = it * G + g;
            const bool live = b < nb;
            if (!live) b = nb - 1;            // keep both halves of the warp convergent; the duplicate's results are dropped
            const int blk = b0 + b;
            const int vlen = min(p.V, p.N - blk * p.V);
            const unsigned vm = vlen == p.V ? vm_full : fs256_valid_mask(p.Lpos, vlen, t);
            float2 xb[16];
            {
                const float4* __restrict__ xs = p.xbs + (size_t)blk * 128 + t;
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) {
                    const float4 q = __ldg(&xs[rr * 16]);
                    xb[2 * rr] = make_float2(q.x, q.y);
                    xb[2 * rr + 1] = make_float2(q.z, q.w);
                }
            }
#pragma unroll 1
            for (int m = 0; m < p.M; ++m) {
                float2 v[16];
                fs256_filter(s_g + (size_t)m * 128 + t, xb, buf, tw, t, v);
                float2 sum2 = make_float2(0.f, 0.f);      // even / odd slots accumulate side by side (one FADD2 per pair)
                float best = 0.f;
#pragma unroll
                for (int s = 0; s < 16; s += 2) {
                    const float m0 = (vm >> s) & 1u ? cabs2(v[s]) : 0.f;
                    const float m1 = (vm >> (s + 1)) & 1u ? cabs2(v[s + 1]) : 0.f;
                    sum2 = __fadd2_rn(sum2, make_float2(m0, m1));
                    best = fmaxf(best, fmaxf(m0, m1));
                }
                acc_sum[m * 17 + t] = sum2.x + sum2.y;   // per-lane partials parked in shared memory: no cross-lane traffic in the loop
                acc_max[m * 17 + t] = best;
            }
            __syncwarp();
            // lane m folds the 16 lanes of mask m in lane order (fixed order -> bit-reproducible) and stores the pair
            for (int m = t; m < p.M; m += 16) {
                float sum = 0.f, best = 0.f;
#pragma unroll
                for (int l = 0; l < 16; ++l) {
                    sum += acc_sum[m * 17 + l];
                    best = fmaxf(best, acc_max[m * 17 + l]);
                }
                if (live) {
                    const size_t o = ((size_t)d * p.nblk + blk) * p.M + m;
                    p.psum[o] = sum;
                    p.pmax[o] = best;
                }
            }
            __syncwarp();                     // the fold has read the partials before the next item overwrites them
        }
        cur += nb;
        // count this CTA's items of the bin; whoever completes the bin reduces and locates it
        __syncthreads();                      // every group's partials are stored
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int prev = atomicAdd(&p.bin_count[d], (unsigned int)nb);
            s_last = (prev + (unsigned int)nb == (unsigned int)p.nblk);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            fs256_finish_bin<G>(p, d, s_g, buf, tw, s_red, s_wblk);
        }
    }
}

// Column reduction of the [nblk][DM] partials in a fixed order, stage 1: slice z of PCS_RED_SLICES takes blocks
// z, z + S, z + 2S, ... and warp w of the CTA every 32nd of those; partial (sum, max, first block) per slice.
#define PCS_RED_SLICES 8
__global__ void __launch_bounds__(1024) search_reduce256_kernel(const float* __restrict__ psum,
                                                                const float* __restrict__ pmax, int DM, int nblk,
                                                                float* __restrict__ part_sum, float* __restrict__ part_max,
                                                                int* __restrict__ part_blk) {
    __shared__ float s_sum[32][33], s_max[32][33];
    __shared__ int s_blk[32][33];
    const int col = blockIdx.x * 32 + threadIdx.x, w = threadIdx.y, z = blockIdx.y;
    float sum = 0.f, best = -1.f;
    int bb = 0x7fffffff;
    if (col < DM) {
        constexpr int STEP = PCS_RED_SLICES * 32;
        int b = z + PCS_RED_SLICES * w;
        for (; b + 3 * STEP < nblk; b += 4 * STEP) {        // four independent loads in flight, same summation order
            float s4[4], v4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const size_t o = (size_t)(b + u * STEP) * DM + col;
                s4[u] = __ldg(&psum[o]);
                v4[u] = __ldg(&pmax[o]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                sum += s4[u];
                if (v4[u] > best) { best = v4[u]; bb = b + u * STEP; }
            }
        }
        for (; b < nblk; b += STEP) {
            const size_t o = (size_t)b * DM + col;
            sum += psum[o];
            const float v = pmax[o];
            if (v > best) { best = v; bb = b; }
        }
    }
    s_sum[w][threadIdx.x] = sum;
    s_max[w][threadIdx.x] = best;
    s_blk[w][threadIdx.x] = bb;
    __syncthreads();
    if (w == 0 && col < DM) {
        for (int k = 1; k < 32; ++k) {
            sum += s_sum[k][threadIdx.x];
            const float v = s_max[k][threadIdx.x];
            const int b = s_blk[k][threadIdx.x];
            if (v > best || (v == best && b < bb)) { best = v; bb = b; }
        }
        part_sum[(size_t)z * DM + col] = sum;
        part_max[(size_t)z * DM + col] = best;
        part_blk[(size_t)z * DM + col] = bb;
    }
}

// Offset of the maximum inside the winning block of every (bin, mask): lowest sample index wins ties.
// Also stage 2 of the reduction: combines the PCS_RED_SLICES partials of its (bin, mask) in slice order and
// writes E (kern:442 scale), the peak value and the winning block.
__global__ void __launch_bounds__(256, 2) peak_locate256_kernel(Os256Params p, const float* __restrict__ part_sum,
                                                                const float* __restrict__ part_max,
                                                                const int* __restrict__ part_blk,
                                                                float* __restrict__ Efull, float* __restrict__ peak_val,
                                                                int* __restrict__ peak_off, unsigned int* done_counter,
                                                                unsigned long long* arrival_flag,
                                                                unsigned long long arrival_value) {
    __shared__ __align__(16) float2 sbuf[16][272];
    const int t = threadIdx.x & 15, g = threadIdx.x >> 4;
    float2* buf = sbuf[g];
    float2 tw[16];
#pragma unroll
    for (int r = 1; r < 16; ++r) tw[r] = __ldg(&p.tw[(t * r) & 255]);
    tw[0] = make_float2(1.f, 0.f);
    const int DM = p.D * p.M;
    int col = blockIdx.x * 16 + g;
    const bool live = col < DM;
    if (!live) col = DM - 1;
    const int d = col / p.M, m = col % p.M;
    int wblk;
    {
        float sum = part_sum[col], pk = part_max[col];
        wblk = part_blk[col];
#pragma unroll
        for (int z = 1; z < PCS_RED_SLICES; ++z) {
            sum += part_sum[(size_t)z * DM + col];
            const float v = part_max[(size_t)z * DM + col];
            const int b = part_blk[(size_t)z * DM + col];
            if (v > pk || (v == pk && b < wblk)) { pk = v; wblk = b; }
        }
        if (wblk == 0x7fffffff) wblk = 0;
        if (live && t == 0) {
            Efull[col] = sum * (1.0f / 262144.0f);      // kern:442 (exact power-of-two scale)
            peak_val[col] = pk;
        }
    }
    const Os256Item it = os256_item(p, (long long)wblk * p.D + d);
    float2 xb[16], v[16];
    os256_block_spectrum(p, it, buf, tw, t, xb);
    os256_filter(p, m, xb, buf, tw, t, v);
    float best = -1.f;
    int idx = 0x7fffffff;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const int rel = t + 16 * dft_q<16>(s) - p.Lpos;
        if (rel >= 0 && rel < it.vlen) {
            const float mag = cabs2(v[s]);
            const int n = it.n0 + rel;
            if (mag > best || (mag == best && n < idx)) { best = mag; idx = n; }
        }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
    }
    if (live && t == 0) peak_off[col] = idx == 0x7fffffff ? 0 : idx;
    // Bin sharding over peer memory: the three tables above may live in another GPU's exchange region.  The last CTA
    // to finish publishes this rank's arrival there (release at system scope after every CTA's stores are fenced).
    if (arrival_flag != nullptr) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int prev = atomicAdd(done_counter, 1u);
            if (prev == gridDim.x - 1) {
                *done_counter = 0;
                __threadfence_system();
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(arrival_flag), "l"(arrival_value) : "memory");
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Reduce the per-block partials (fixed order -> bit-reproducible, unlike the reference's float
// atomics, kern:463,474).  One warp per (d, m): E, the largest |y|^2 and the block that holds it (lowest block on ties);
// search_os_kernel<.., LOCATE = true> then turns that block into the peak's sample offset.
// ---------------------------------------------------------------------------------------------
__global__ void search_reduce_kernel(const float* __restrict__ psum, const float* __restrict__ pmax, int DM, int nblk,
                                     float* __restrict__ Efull, float* __restrict__ peak_val, int* __restrict__ wblk) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= DM) return;
    PeakAcc acc;
    acc.init();
    const size_t base = (size_t)w * nblk;
    for (int b = lane; b < nblk; b += 32) {
        acc.sum += psum[base + b];
        acc.merge(pmax[base + b], b);
    }
    acc.warp_reduce();
    if (lane == 0) {
        Efull[w] = acc.sum * (1.0f / 262144.0f);      // kern:442 (exact power-of-two scale)
        peak_val[w] = acc.best;
        wblk[w] = acc.idx == 0x7fffffff ? 0 : acc.idx;
    }
}

// ---------------------------------------------------------------------------------------------
// a9 findDopplerEst (kern:502-597) + the host interpolation of dem_base:610-618 done on the
// device in float64 so that the demod stage can be enqueued without a host round trip, + the
// north-star peak, + the spectrum windows computeSNR needs (dem_base:635-667).
// Single CTA.
// ---------------------------------------------------------------------------------------------
struct EstimateParams {
    const float* __restrict__ Efull;    // [D][M] per-mask energies
    float* __restrict__ E;              // [D][M] reference layout (SUM mode: column 0 only)
    const float* __restrict__ peak_val; // [D][M]
    const int* __restrict__ peak_off;   // [D][M]
    const int* __restrict__ shifts;     // [D]
    DevResult* __restrict__ res;
    int D, M, N, num_dopplers, element_offset, sum_all, window_width, window_cap;
};

__global__ void __launch_bounds__(256) estimate_kernel(EstimateParams p) {
    __shared__ float s_idx[32], s_val[32];
    __shared__ float s_pk[256];
    __shared__ int s_pki[256];
    const int tid = threadIdx.x;
    // 1 + 2. reference layout of E and the per-column running top-2 (kern:527-544, one thread per column), staged
    // through shared memory 256 bins at a time so that the serial scan runs at shared-memory latency.
    __shared__ float tile[256][33];
    const int ncols = p.sum_all ? 1 : p.M;
    float v0 = 0.f, v1 = 0.f;
    int i0 = 0, i1 = 0, cur = 0;
    for (int base = 0; base < p.D; base += 256) {
        const int rows = min(256, p.D - base);
        if (p.sum_all) {
            if (tid < rows) {
                const float* src = p.Efull + (size_t)(base + tid) * p.M;
                float acc = src[0];
                for (int m = 1; m < p.M; ++m) acc += src[m];                          // kern:459-462
                float* dst = p.E + (size_t)(base + tid) * p.M;
                dst[0] = acc;
                for (int m = 1; m < p.M; ++m) dst[m] = 0.f;
                tile[tid][0] = acc;
            }
        } else {
            for (int i = tid; i < rows * p.M; i += blockDim.x) {
                const float e = p.Efull[(size_t)base * p.M + i];
                p.E[(size_t)base * p.M + i] = e;
                tile[i / p.M][i % p.M] = e;
            }
        }
        __syncthreads();
        if (tid < ncols) {
            const int lo = max(base, p.element_offset), hi = min(base + rows, p.num_dopplers + p.element_offset);
            for (int i = lo; i < hi; ++i) {
                const float e = tile[i - base][tid];
                if (e > (cur ? v1 : v0)) {
                    if (cur) { v1 = e; i1 = i; } else { v0 = e; i0 = i; }
                    cur = (v0 >= v1) ? 1 : 0;
                }
            }
        }
        __syncthreads();
    }
    if (tid < ncols) {
        const float tmp = __fmaf_rn((float)i0, v0, __fmul_rn((float)i1, v1));   // SASS: FMUL + FFMA
        float bl = __fdiv_rn(tmp, __fadd_rn(v0, v1));
        float vl = __fdiv_rn(tmp, (float)(i0 + i1));
        if (p.element_offset > 0) vl = __fdiv_rn(cur ? v0 : v1, p.E[tid]);    // kern:550-554
        s_idx[tid] = bl;
        s_val[tid] = vl;
    }
    __syncthreads();
    if (tid == 0) {
        float best, metric;
        if (p.sum_all) {
            best = s_idx[0];
            metric = s_val[0];
        } else {
            // mean over masks (xor-butterfly order for power-of-two M, kern:576-580)
            float a[32], b[32];
            for (int m = 0; m < p.M; ++m) { a[m] = s_idx[m]; b[m] = s_val[m]; }
            if ((p.M & (p.M - 1)) == 0) {
                for (int step = p.M >> 1; step > 0; step >>= 1)
                    for (int m = 0; m < step; ++m) { a[m] = a[m] + a[m ^ step]; b[m] = b[m] + b[m ^ step]; }
            } else {
                for (int m = 1; m < p.M; ++m) { a[0] += a[m]; b[0] += b[m]; }
            }
            best = a[0] / (float)p.M;
            metric = b[0] / (float)p.M;
        }
        DevResult* r = p.res;
        r->best_idx = best;
        r->metric_db = 10.f * log10f(metric);
        // host interpolation, float64 (dem_base:610-618)
        if (isnan(best) || best < 0.f || best > (float)(p.D - 1)) {
            r->status = 1;
            r->low_idx = r->high_idx = 0;
            r->shift = 0;
        } else {
            const double b = (double)best;
            const int lo = (int)b, hi = (int)ceil(b);
            const double frac = fmod(b, 1.0);
            const int slo = p.shifts[lo], shi = p.shifts[hi];
            r->status = 0;
            r->low_idx = lo;
            r->high_idx = hi;
            r->shift = (int)rint((double)slo + (double)(shi - slo) * frac);   // np.round: half to even
        }
    }
    // 3. north-star peak over (bin, mask): lowest flat index wins ties
    {
        float bv = -1.f;
        int bi = 0x7fffffff;
        const int first = p.element_offset * p.M, last = (p.element_offset + p.num_dopplers) * p.M;
        for (int i = first + tid; i < last; i += blockDim.x) {
            const float v = p.peak_val[i];
            if (v > bv) { bv = v; bi = i; }
        }
        s_pk[tid] = bv;
        s_pki[tid] = bi;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (tid < s) {
                const float ov = s_pk[tid + s];
                const int oi = s_pki[tid + s];
                if (ov > s_pk[tid] || (ov == s_pk[tid] && oi < s_pki[tid])) { s_pk[tid] = ov; s_pki[tid] = oi; }
            }
            __syncthreads();
        }
        if (tid == 0) {
            const int i = s_pki[0];
            const bool have = i != 0x7fffffff;            // false for the Parseval variant (no surface, no peak)
            p.res->peak_val = have ? s_pk[0] : -1.f;
            p.res->peak_bin = have ? i / p.M : -1;
            p.res->peak_mask = have ? i % p.M : -1;
            p.res->peak_offset = have ? p.peak_off[i] : -1;
        }
    }
    __syncthreads();
    // 4. geometry of the spectrum windows computeSNR averages: circular [s_lo - w, s_hi + w) and the same + N/2
    //    (the bins themselves are produced by spectrum_bins_kernel)
    if (tid == 0) {
        DevResult* r = p.res;
        int sig_start = 0, len = 0, noise_start = 0;
        if (r->status == 0) {
            const int slo = p.shifts[r->low_idx], shi = p.shifts[r->high_idx];
            const int nmask = p.N - 1;
            sig_start = (slo - p.window_width) & nmask;
            len = ((shi - slo) & nmask) + 2 * p.window_width;
            if (len > p.window_cap) len = 0;      // host falls back to a full spectrum read
            noise_start = (sig_start + p.N / 2) & nmask;
        }
        r->sig_start = sig_start;
        r->sig_len = len;
        r->noise_start = noise_start;
        r->noise_len = len;
    }
}

// ---------------------------------------------------------------------------------------------
// a12 + a13 (first half) fused, overlap-save form: y[m, n] at the selected shift for all masks;
// writes |y|^2 (what findCentres consumes, kern:111,129) and p[n] = sum_m |y|^2 (kern:191-205);
// optionally the complex surface for inspection.  The shift comes from device memory.
// ---------------------------------------------------------------------------------------------
struct OsDemodParams {
    const float2* __restrict__ x;
    const float2* __restrict__ gb;
    const float2* __restrict__ tw;
    const DevResult* __restrict__ res;   // shift source when shift_override < 0
    float* __restrict__ ymag;            // [M][N]
    float* __restrict__ p;               // [N]
    float2* __restrict__ ycplx;          // [M][N] or null
    int N, M, nblk, V, Lpos, shift_override, mask_lo, mask_hi;
    float invN;
};

template <int LOGB, int G>
__global__ void __launch_bounds__(G * FftShape<LOGB>::T) demod_os_kernel(OsDemodParams p) {
    using S = FftShape<LOGB>;
    constexpr int B = S::B, T = S::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* smem = reinterpret_cast<float2*>(smem_raw);
    const int g = threadIdx.x / T, t = threadIdx.x % T;
    float2* xb = smem + (size_t)g * 3 * S::WORK;
    float2* work0 = xb + S::WORK;
    float2* work1 = work0 + S::WORK;
    const int bar_id = 1 + g;
    const int blk = blockIdx.x * G + g;
    if (blk >= p.nblk) return;
    const uint32_t nmask = (uint32_t)p.N - 1u;
    const uint32_t shift = (uint32_t)(p.shift_override >= 0 ? p.shift_override : p.res->shift) & nmask;
    const int n0 = blk * p.V;
    const uint32_t n_first = (uint32_t)(n0 - p.Lpos) & nmask;
    {
        RotatePre pre;
        pre.shift = shift;
        pre.n_first = n_first;
        pre.nmask = nmask;
        pre.invN = p.invN;
        pre.step = unit_phasor_neg((shift * (uint32_t)(B / 16)) & nmask, p.invN);
        auto src = [&](int i) { return __ldg(&p.x[(n_first + (uint32_t)i) & nmask]); };
        auto sink = [&](int i, float2 v, int) { xb[padi(i)] = v; };
        group_fft<LOGB, -1>(work0, work1, p.tw, t, bar_id, src, sink, pre);
    }
    group_sync<T>(bar_id);
    const int vlen = min(p.V, p.N - n0);
    float psum[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) psum[s] = 0.f;
    for (int m = 0; m < p.M; ++m) {
        const float2* __restrict__ gm = p.gb + (size_t)m * B;
        const bool in_sum = (m >= p.mask_lo && m < p.mask_hi);
        auto src = [&](int i) { return cmul(xb[padi(i)], __ldg(&gm[i])); };
        auto sink = [&](int i, float2 v, int slot) {
            const int rel = i - p.Lpos;
            const float mag = cabs2(v);
            if (in_sum) psum[slot] += mag;
            if (rel >= 0 && rel < vlen) {
                const size_t o = (size_t)m * p.N + n0 + rel;
                p.ymag[o] = mag;
                if (p.ycplx) p.ycplx[o] = v;
            }
        };
        group_fft<LOGB, +1>(work0, work1, p.tw, t, bar_id, src, sink);
    }
    // p[n]: replay the last pass's index map (slot -> output index) without data
    {
        constexpr int R = S::RLAST, NB = 16 / R, NS = B / R;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int j = t + b * T;
#pragma unroll
            for (int s = 0; s < R; ++s) {
                const int i = j + dft_q<R>(s) * NS;
                const int rel = i - p.Lpos;
                if (rel >= 0 && rel < vlen) p.p[n0 + rel] = psum[b * R + s];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// a12 + a13 (first half), register-resident 256-point form: one group per block, all masks.
// ---------------------------------------------------------------------------------------------
struct Demod256Params {
    Os256Params os;                      // x, gperm, tw, N, M, nblk, V, Lpos, invN (D / shifts / psum / pmax unused)
    const DevResult* __restrict__ res;   // shift source when shift_override < 0
    float* __restrict__ ymag;            // [M][N]
    float* __restrict__ p;               // [N]
    float2* __restrict__ ycplx;          // [M][N] or null
    int shift_override, mask_lo, mask_hi;
};

__global__ void __launch_bounds__(64) demod_os256_kernel(Demod256Params q) {
    __shared__ __align__(16) float2 sbuf[4][272];
    const Os256Params& p = q.os;
    const int t = threadIdx.x & 15, g = threadIdx.x >> 4;
    float2* buf = sbuf[g];
    float2 tw[16];
#pragma unroll
    for (int r = 1; r < 16; ++r) tw[r] = __ldg(&p.tw[(t * r) & 255]);
    tw[0] = make_float2(1.f, 0.f);
    int blk = blockIdx.x * 4 + g;
    const bool live = blk < p.nblk;
    if (!live) blk = p.nblk - 1;
    Os256Item it;
    it.blk = blk;
    it.d = 0;
    it.nmask = (uint32_t)p.N - 1u;
    it.shift = (uint32_t)(q.shift_override >= 0 ? q.shift_override : q.res->shift) & it.nmask;
    it.n0 = blk * p.V;
    it.n_first = (uint32_t)(it.n0 - p.Lpos) & it.nmask;
    it.vlen = min(p.V, p.N - it.n0);
    float2 xb[16];
    os256_block_spectrum(p, it, buf, tw, t, xb);
    float psum[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) psum[s] = 0.f;
    for (int m = 0; m < p.M; ++m) {
        float2 v[16];
        os256_filter(p, m, xb, buf, tw, t, v);
        const bool in_sum = (m >= q.mask_lo && m < q.mask_hi);
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            const float mag = cabs2(v[s]);
            if (in_sum) psum[s] += mag;                       // mask order, like kern:197-201
            const int rel = t + 16 * dft_q<16>(s) - p.Lpos;
            if (live && rel >= 0 && rel < it.vlen) {
                const size_t o = (size_t)m * p.N + it.n0 + rel;
                q.ymag[o] = mag;
                if (q.ycplx) q.ycplx[o] = v[s];
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const int rel = t + 16 * dft_q<16>(s) - p.Lpos;
        if (live && rel >= 0 && rel < it.vlen) q.p[it.n0 + rel] = psum[s];
    }
}

// ---------------------------------------------------------------------------------------------
// Pruned second pass of the four-step transform.  After pass 1 the scratch holds S[k1][n2] (twiddled); output bin
// k = k1 + N1 k2 is sum_n2 S[k1][n2] W_N2^(n2 k2).  Timing recovery only looks at the symbol-rate band
// [iH, iL) (dem_base:508-512), i.e. nk2 <= 16 consecutive k2 for every k1: one warp per k1 row, direct sums.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) band_dft_kernel(const float2* __restrict__ S, const float2* __restrict__ tw2,
                                                       int N1, int N2, int k2lo, int nk2, float2* __restrict__ out) {
    const int k1 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k1 >= N1) return;
    float2 acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = make_float2(0.f, 0.f);
    const float2* row = S + (size_t)k1 * N2;
    for (int n2 = lane; n2 < N2; n2 += 32) {
        const float2 v = row[n2];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (j < nk2) {
                const float2 w = __ldg(&tw2[(n2 * (k2lo + j)) & (N2 - 1)]);
                acc[j].x = fmaf(v.x, w.x, fmaf(-v.y, w.y, acc[j].x));
                acc[j].y = fmaf(v.x, w.y, fmaf(v.y, w.x, acc[j].y));
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (j < nk2) {
            float2 a = acc[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
                a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
            }
            if (lane == 0) out[(size_t)k1 + (size_t)N1 * (k2lo + j)] = a;
        }
    }
}

// Same idea for computeSNR (dem_base:635-667): the two circular spectrum windows the estimate selected, one warp
// per bin.  Window geometry comes from the DevResult the estimate kernel just wrote.
__global__ void __launch_bounds__(256) spectrum_bins_kernel(const float2* __restrict__ S, const float2* __restrict__ tw2,
                                                            int N1, int N2, const DevResult* __restrict__ res,
                                                            float2* __restrict__ sig_win, float2* __restrict__ noise_win) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int len = res->sig_len;
    if (w >= 2 * len) return;
    const int which = w >= len, i = which ? w - len : w;
    const int N = N1 * N2;
    const int k = ((which ? res->noise_start : res->sig_start) + i) & (N - 1);
    const int k1 = k & (N1 - 1), k2 = k / N1;
    const float2* row = S + (size_t)k1 * N2;
    float2 a = make_float2(0.f, 0.f);
    for (int n2 = lane; n2 < N2; n2 += 32) {
        const float2 v = row[n2];
        const float2 wv = __ldg(&tw2[(n2 * k2) & (N2 - 1)]);
        a.x = fmaf(v.x, wv.x, fmaf(-v.y, wv.y, a.x));
        a.y = fmaf(v.x, wv.y, fmaf(v.y, wv.x, a.y));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
        a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
    }
    if (lane == 0) (which ? noise_win : sig_win)[i] = a;
}

// ---------------------------------------------------------------------------------------------
// Large 1-D FFT as two tiled passes (N = N1 * N2, "four-step"): each CTA transforms C vectors
// of length B = 2^LOGB that are strided in global memory, staged through a shared tile so both
// the load and the store are coalesced.  Used for the chunk spectrum X (a4, dem_base:557), the
// timing-recovery spectrum (a13, dem_base:721) and the full-length search/demod path.
//   value loaded for (vector v, element e):  load(v * in_vs + e * in_es)
//   value stored for (vector v, element e):  store(v * out_vs + e * out_es, val)
//   optional twiddle after the transform:    val *= exp(DIR * 2 pi i * v * e / Ntw)
// ---------------------------------------------------------------------------------------------
struct TileGeom {
    int nvec;
    long long in_vs, in_es, out_vs, out_es;
    int v_fast_in, v_fast_out;   // 1: consecutive vectors are adjacent in memory (coalesce over v)
    int twiddle;                 // apply inter-pass twiddle
    float inv_ntw;
    uint32_t ntw_mask;
};

template <int LOGB, int DIR, int C, typename Load, typename Store>
__global__ void __launch_bounds__(C * FftShape<LOGB>::T) fft_tile_kernel(TileGeom g, const float2* __restrict__ tw,
                                                                          Load load, Store store) {
    using S = FftShape<LOGB>;
    constexpr int B = S::B, T = S::T, NT = C * T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* bufA = reinterpret_cast<float2*>(smem_raw);   // [C][WORK]
    float2* bufB = bufA + (size_t)C * S::WORK;            // [C][WORK]
    const int v0 = blockIdx.x * C;
    const int tid = threadIdx.x;
    // ---- coalesced tile load ----
    for (int i = tid; i < C * B; i += NT) {
        int v, e;
        if (g.v_fast_in) { v = i % C; e = i / C; } else { e = i % B; v = i / B; }
        float2 val = make_float2(0.f, 0.f);
        if (v0 + v < g.nvec) val = load((long long)(v0 + v) * g.in_vs + (long long)e * g.in_es);
        bufA[(size_t)v * S::WORK + padi(e)] = val;
    }
    __syncthreads();
    // ---- one group per vector ----
    {
        const int v = tid / T, t = tid % T;
        float2* a = bufA + (size_t)v * S::WORK;
        float2* b = bufB + (size_t)v * S::WORK;
        // pass 0 reads a -> b, pass 1 b -> a, pass 2 a -> ...: final sink goes to the buffer the last
        // pass does not read (b for odd NPASS, a for even NPASS).
        float2* outbuf = (S::NPASS & 1) ? b : a;
        const uint32_t vg = (uint32_t)(v0 + v);
        auto src = [&](int i) { return a[padi(i)]; };
        auto sink = [&](int i, float2 val, int) {
            if (g.twiddle) {
                float2 w = unit_phasor_neg((vg * (uint32_t)i) & g.ntw_mask, g.inv_ntw);
                if (DIR > 0) w = cconj(w);
                val = cmul(val, w);
            }
            outbuf[padi(i)] = val;
        };
        group_fft<LOGB, DIR>(b, a, tw, t, 1 + (v % 15), src, sink);
    }
    __syncthreads();
    // ---- coalesced tile store ----
    float2* outb = (S::NPASS & 1) ? bufB : bufA;
    for (int i = tid; i < C * B; i += NT) {
        int v, e;
        if (g.v_fast_out) { v = i % C; e = i / C; } else { e = i % B; v = i / B; }
        if (v0 + v < g.nvec)
            store((long long)(v0 + v) * g.out_vs + (long long)e * g.out_es, outb[(size_t)v * S::WORK + padi(e)]);
    }
}

// ---------------------------------------------------------------------------------------------
// a14 findCodeRateAndPhase (kern:236-320) + the host arithmetic of dem_base:733-750 (float64).
// Single CTA; lowest index wins ties.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) timing_kernel(const float2* __restrict__ Pf, int offset, int len, int N,
                                                     int spsym_min, DevResult* __restrict__ res) {
    __shared__ float s_v[32];
    __shared__ int s_i[32];
    float bv = -1.f;
    int bi = 0x7fffffff;
    for (int x = threadIdx.x; x < len; x += blockDim.x) {
        const float v = cabs2(Pf[offset + x]);
        if (v > bv) { bv = v; bi = offset + x; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = bv; s_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (s_v[w] > bv || (s_v[w] == bv && s_i[w] < bi)) { bv = s_v[w]; bi = s_i[w]; }
        if (bi == 0x7fffffff) bi = offset;
        const float2 z = Pf[bi];
        const float phase = atan2f(z.y, z.x);
        res->timing[0] = (float)bi;
        res->timing[1] = phase;
        res->timing[2] = cabs2(z);
        double sp = (double)N / (double)(float)bi;                     // dem_base:735
        double off = -(double)phase / 3.141592653589793 * sp / 2.0;    // dem_base:745
        if (off < 0) off += sp - 1.0;                                  // dem_base:746-747
        res->sp_sym = sp;
        res->code_offset = off;
        double spc = sp < (double)spsym_min ? (double)spsym_min : sp;  // dem_base:994-995
        res->n_sym = (int)((double)N / spc);                           // dem_base:999
    }
}

// ---------------------------------------------------------------------------------------------
// a15 findCentres (kern:78-146), CENTRES_ABS only (the reference never passes another op,
// dem_base:796).  One thread per symbol; bit-exact index arithmetic incl. the FFMA contraction.
// ---------------------------------------------------------------------------------------------
__global__ void centres_kernel(const float* __restrict__ ymag, const DevResult* __restrict__ res, int N, int M, int W,
                               int spsym_min, int max_sym, int* __restrict__ out_sym, int* __restrict__ out_idx,
                               float* __restrict__ out_mag) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const double spd = res->sp_sym < (double)spsym_min ? (double)spsym_min : res->sp_sym;
    const int n_sym = min(res->n_sym, max_sym);
    if (x >= n_sym) return;
    const float spSym = (float)spd;                 // np.float32(spSym), dem_base:997
    const float offset = (float)res->code_offset;   // np.float32(codePhase)
    const int left = W / 2;
    const float tbase = __fmaf_rn((float)x, spSym, -(float)left);   // FFMA R, R, -3 in the reference SASS
    int arrayIdx = (int)__fadd_rn(tbase, offset);
    int maxArrayIdx = arrayIdx + W;
    int offsetComp = (int)offset;
    if (arrayIdx < 0) {
        offsetComp -= arrayIdx;
        arrayIdx = 0;
    }
    if (maxArrayIdx > N) maxArrayIdx = N;
    maxArrayIdx -= arrayIdx;
    int maxIdx = -1, maxCentreIdx = -1;
    float maxVal = 0.f;
    if (arrayIdx < N) {
        for (int m = 0; m < M; ++m) {
            const float* row = ymag + (size_t)m * N + arrayIdx;
            for (int k = 0; k < maxArrayIdx; ++k) {
                const float v = row[k];
                if (v > maxVal) {
                    maxVal = v;
                    maxIdx = m;
                    maxCentreIdx = k;
                }
            }
        }
        out_sym[x] = maxIdx;
        out_idx[x] = (int)__fadd_rn(__fadd_rn(tbase, (float)maxCentreIdx), (float)offsetComp);
        out_mag[x] = maxVal;
    }
}

// ---------------------------------------------------------------------------------------------
// Parseval variant of the search metric -- a LABELLED alternative, never the default:
//   E[d,m] = sum_n |y[d,m,n]|^2 / 2^18 = (N / 2^18) * sum_k |X[(k+s_d) % N]|^2 * |Mk[m,k]|^2
// needs no inverse transform at all (and yields no peak / offsets).  W holds |Mk|^2 (or, in SUM mode, its sum over
// the masks: one row).  One warp per (256-bin spectrum tile, Doppler-bin slice): the lane keeps its 8 x MW weights
// in registers and walks the Doppler bins; partials [tile][D][MW] are reduced in a fixed order afterwards.
// ---------------------------------------------------------------------------------------------
__global__ void abs2_kernel(const float2* __restrict__ X, float* __restrict__ PX, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) PX[i] = cabs2(X[i]);
}

template <int MW>
__global__ void __launch_bounds__(256) parseval_energy_kernel(const float* __restrict__ PX, const float* __restrict__ W,
                                                              const int* __restrict__ shifts, float* __restrict__ part,
                                                              int N, int D, int d_per_warp) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int ntile = N >> 8;
    const int tile = warp % ntile, dslice = warp / ntile;
    const int d0 = dslice * d_per_warp, d1 = min(D, d0 + d_per_warp);
    if (d0 >= D) return;
    const int k0 = tile << 8;
    float w[8][MW];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int m = 0; m < MW; ++m) w[j][m] = __ldg(&W[(size_t)m * N + k0 + lane + 32 * j]);
    const uint32_t nmask = (uint32_t)N - 1u;
    for (int d = d0; d < d1; ++d) {
        const uint32_t s = (uint32_t)shifts[d];
        float acc[MW];
#pragma unroll
        for (int m = 0; m < MW; ++m) acc[m] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float px = __ldg(&PX[((uint32_t)(k0 + lane + 32 * j) + s) & nmask]);
#pragma unroll
            for (int m = 0; m < MW; ++m) acc[m] = fmaf(px, w[j][m], acc[m]);
        }
#pragma unroll
        for (int m = 0; m < MW; ++m) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int m = 0; m < MW; ++m) part[((size_t)tile * D + d) * MW + m] = acc[m];
        }
    }
}

// Fixed-order sum over the tiles; scale N / 2^18.  Writes columns [0, MW) of rows with stride M (the pointers are
// already offset to the batch's first column) and clears `extra` further columns (SUM mode: only column 0 is used).
__global__ void __launch_bounds__(1024) parseval_reduce_kernel(const float* __restrict__ part, int ntile, int D, int MW,
                                                               int M, int extra, float scale, float* __restrict__ Efull,
                                                               float* __restrict__ peak_val, int* __restrict__ peak_off) {
    // block (32, 32): lane = column (d, mw), warp w sums tiles w, w + 32, ...; the 32 partials are added in warp order
    __shared__ float s_sum[32][33];
    const int col = blockIdx.x * 32 + threadIdx.x, w = threadIdx.y, cols = D * MW;
    float s0 = 0.f, s1 = 0.f;
    if (col < cols) {
        int t = w;
        for (; t + 32 < ntile; t += 64) {
            s0 += part[(size_t)t * cols + col];
            s1 += part[(size_t)(t + 32) * cols + col];
        }
        for (; t < ntile; t += 32) s0 += part[(size_t)t * cols + col];
    }
    s_sum[w][threadIdx.x] = s0 + s1;
    __syncthreads();
    if (w != 0 || col >= cols) return;
    float sum = s_sum[0][threadIdx.x];
    for (int k = 1; k < 32; ++k) sum += s_sum[k][threadIdx.x];
    const int d = col / MW, m = col % MW;
    Efull[(size_t)d * M + m] = sum * scale;
    peak_val[(size_t)d * M + m] = -1.f;        // this variant has no correlation surface, hence no peak
    peak_off[(size_t)d * M + m] = -1;
    if (m == MW - 1)
        for (int mm = MW; mm < MW + extra; ++mm) {
            Efull[(size_t)d * M + mm] = 0.f;
            peak_val[(size_t)d * M + mm] = -1.f;
            peak_off[(size_t)d * M + mm] = -1;
        }
}

// ---------------------------------------------------------------------------------------------
// a19 __thresholdInput (dem_base:670-707) on the device: two passes of "clip everything above
// peakThresholdScale * mean(|x|) back onto that radius".  The mean is np.mean's: float32 pairwise summation in NumPy's
// own order (leaves of 128 samples, eight strided accumulators per leaf combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)),
// then a balanced tree over the leaves -- loops_utils.h.src: pairwise_sum), so for a power-of-two chunk the threshold
// is the float32 NumPy computes from the same magnitudes.  The clip is thresh * (x / |x|) with NumPy's complex
// arithmetic (x * fl(1/|x|), then * thresh, every operation rounded to float32).
// ---------------------------------------------------------------------------------------------
// Sum of 1024 floats in shared memory in NumPy's pairwise order; 128 threads; result valid in thread 0.
PCS_DEVINL float np_pairwise_sum_1024(const float* s, float* scratch) {
    float r = 0.f;
    const int tid = threadIdx.x;
    if (tid < 64) {
        const float* leaf = s + (tid >> 3) * 128 + (tid & 7);
        r = leaf[0];
#pragma unroll
        for (int i = 1; i < 16; ++i) r = __fadd_rn(r, leaf[8 * i]);
#pragma unroll
        for (int o = 1; o <= 16; o <<= 1) r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, o));   // within leaf, then leaves 2 by 2
        if (tid == 32) scratch[0] = r;
    }
    __syncthreads();
    if (tid == 0) r = __fadd_rn(r, scratch[0]);
    return r;
}

// pass 0: |x| and the per-1024 partial sums.
__global__ void __launch_bounds__(128) threshold_abs_kernel(const float2* __restrict__ x, float* __restrict__ mag,
                                                            float* __restrict__ partial) {
    __shared__ float s[1024];
    __shared__ float scratch[1];
    const size_t base = (size_t)blockIdx.x * 1024;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float2 v = x[base + threadIdx.x + 128 * k];
        const float m = hypotf(v.x, v.y);
        s[threadIdx.x + 128 * k] = m;
        mag[base + threadIdx.x + 128 * k] = m;
    }
    __syncthreads();
    const float r = np_pairwise_sum_1024(s, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

// thresh = float32(scale) * (pairwise total / N): balanced tree over nb (power of two, <= 4096) partials.
__global__ void __launch_bounds__(1024) threshold_level_kernel(const float* __restrict__ partial, int nb, float scale, float invN,
                                                               float* __restrict__ thresh) {
    __shared__ float s[4096];
    for (int i = threadIdx.x; i < nb; i += 1024) s[i] = partial[i];
    __syncthreads();
    for (int stride = 1; stride < nb; stride <<= 1) {
        for (int i = threadIdx.x * 2 * stride; i + stride < nb; i += 2048 * stride) s[i] = __fadd_rn(s[i], s[i + stride]);
        __syncthreads();
    }
    if (threadIdx.x == 0) *thresh = __fmul_rn(scale, __fmul_rn(__fadd_rn(0.f, s[0]), invN));
}

// pass 1 (LAST = false): clip, refresh |x| of the clipped samples, partial sums of the refreshed magnitudes;
// pass 2 (LAST = true): clip and mark the clipped samples in a bit mask (clippedPeakIPure).
template <bool LAST>
__global__ void __launch_bounds__(128) threshold_clip_kernel(float2* __restrict__ x, float* __restrict__ mag,
                                                             const float* __restrict__ thresh, float* __restrict__ partial,
                                                             unsigned int* __restrict__ bits) {
    __shared__ float s[1024];
    __shared__ float scratch[1];
    const float thr = *thresh;
    const size_t base = (size_t)blockIdx.x * 1024;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const size_t n = base + threadIdx.x + 128 * k;
        float m = mag[n];
        const bool over = m > thr;
        if (over) {
            const float2 v = x[n];
            // NumPy's complex64 / (m + 0i): rat = 0 / m, scl = 1 / (m + 0 * rat), q = ((re + im rat) scl, (im - re rat) scl);
            // then (thr + 0i) * q = (thr q.re - 0 q.im, thr q.im + 0 q.re).  Spelled out term by term so that signed
            // zeros and non-finite samples come out as they do there.
            const float rat = __fdiv_rn(0.f, m);
            const float scl = __frcp_rn(__fadd_rn(m, __fmul_rn(0.f, rat)));
            const float qr = __fmul_rn(__fadd_rn(v.x, __fmul_rn(v.y, rat)), scl);
            const float qi = __fmul_rn(__fsub_rn(v.y, __fmul_rn(v.x, rat)), scl);
            const float2 w = make_float2(__fsub_rn(__fmul_rn(thr, qr), __fmul_rn(0.f, qi)),
                                         __fadd_rn(__fmul_rn(thr, qi), __fmul_rn(0.f, qr)));
            x[n] = w;
            if (!LAST) {
                m = hypotf(w.x, w.y);
                mag[n] = m;
            }
        }
        if (LAST) {
            const unsigned int word = __ballot_sync(0xffffffffu, over);
            if ((threadIdx.x & 31) == 0) bits[n >> 5] = word;
        } else {
            s[threadIdx.x + 128 * k] = m;
        }
    }
    if (!LAST) {
        __syncthreads();
        const float r = np_pairwise_sum_1024(s, scratch);
        if (threadIdx.x == 0) partial[blockIdx.x] = r;
    }
}

// ---------------------------------------------------------------------------------------------
// Peer exchange (bin sharding over NVLink): monotonic 64-bit sequence flags in the ranks' exchange regions, written with
// system-scope releases by kernels of one GPU and acquired by tiny wait kernels of another.  A wait gives up after ~10 s
// (a peer died) and records it instead of hanging the GPU; a flag NEWER than the one waited for where that cannot
// legally happen (row flags: the back-pressure of pcs_shard_submit forbids it) is recorded as an overrun.
// ---------------------------------------------------------------------------------------------
#define PCS_WAIT_CYCLES 20000000000ll
PCS_DEVINL unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
PCS_DEVINL void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// flags[i] = value for up to 32 destinations (possibly on other GPUs), after everything earlier in the stream.
struct FlagList {
    unsigned long long* dst[16];
    int n;
};
// tagp != null (a launch replayed from a CUDA graph, whose parameters are frozen): the value is read from *tagp, which is then
// advanced by inc for the next replay.
__global__ void flags_set_kernel(FlagList f, unsigned long long value, unsigned long long* tagp = nullptr,
                                 unsigned long long inc = 0ull) {
    if (tagp) value = *tagp;
    __threadfence_system();
    if ((int)threadIdx.x < f.n) st_release_sys(f.dst[threadIdx.x], value);
    if (tagp) {
        __syncwarp();
        if (threadIdx.x == 0) *tagp = value + inc;
    }
}

// Lane i < n spins until flags[i * stride] >= want[i] (want 0 = nothing to wait for).  err (device word, may be null)
// gets bit 0 on a timeout.  exact != 0: a flag beyond `want` sets bit 1 (overrun).  res (may be null): xchg_timeout.
struct WaitList {
    const unsigned long long* src[20];
    unsigned long long want[20];
    int n;
};
// tagp != null (graph replay): every non-zero `want` is replaced by *tagp.
__global__ void flags_wait_kernel(WaitList w, int exact, int* err, DevResult* res, const unsigned long long* tagp = nullptr) {
    const int i = threadIdx.x;
    int bad = 0;
    if (i < w.n && w.want[i] != 0ull) {
        const unsigned long long want = tagp ? *tagp : w.want[i];
        const long long t0 = clock64();
        for (;;) {
            const unsigned long long v = ld_acquire_sys(w.src[i]);
            if (v >= want) {
                if (exact && v > want) bad = 2;
                break;
            }
            if (clock64() - t0 > PCS_WAIT_CYCLES) { bad = 1; break; }
            __nanosleep(100);
        }
    }
    const unsigned any = __ballot_sync(0xffffffffu, bad != 0);
    if (bad && err) atomicOr(err, bad);
    if (i == 0 && res) res->xchg_timeout = any ? 1 : 0;
    __threadfence_system();
}

// Ingest rank: chunk `src` -> ring slots (own / peers') and, optionally, the chunk's block spectra `src2` -> the peers' block-
// spectra rings (NVLink P2P stores), as ONE launch:
//   prologue  thread 0 of every CTA waits until the slot is free everywhere (the acks in `w`, see csrc/shard.inc)
//   body      float4 loads of the chunk, one store per destination
//   epilogue  the CTA that finishes last raises every peer's data flag (system-scope release after the fenced stores)
struct BcastParams {
    const float4* __restrict__ src;
    float4* dst[16];            // job 1 (the chunk): destinations; one equal to src is skipped
    int ndst;
    long long n4;               // float4 elements
    const float4* __restrict__ src2;   // job 2 (optional: the chunk's block spectra): same idea
    float4* dst2[16];
    int ndst2;
    long long n4_2;
    WaitList w;
    FlagList f;
    unsigned long long value;
    unsigned int* done;         // CTA completion counter (self-resetting)
    int* err;
};
// 128-thread CTAs at <= 32 registers: two of them fit on an SM NEXT TO a full complement of search CTAs (4 x 128 threads x 112
// registers leave 8192 registers and 58 KB of shared memory), so the broadcast never waits for a search wave to retire.
__global__ void __launch_bounds__(128, 16) chunk_bcast_kernel(BcastParams p) {
    if (threadIdx.x == 0) {
        int bad = 0;
        for (int i = 0; i < p.w.n; ++i) {
            if (p.w.want[i] == 0ull) continue;
            const long long t0 = clock64();
            while (ld_acquire_sys(p.w.src[i]) < p.w.want[i]) {
                if (clock64() - t0 > PCS_WAIT_CYCLES) { bad = 1; break; }
                __nanosleep(100);
            }
        }
        if (bad && p.err) atomicOr(p.err, 1);
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (p.ndst > 0)
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n4; i += stride) {
            const float4 v = __ldg(p.src + i);
#pragma unroll 4
            for (int d = 0; d < p.ndst; ++d)
                if (p.dst[d] != p.src) p.dst[d][i] = v;
        }
    if (p.ndst2 > 0)
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n4_2; i += stride) {
            const float4 v = __ldg(p.src2 + i);
#pragma unroll 4
            for (int d = 0; d < p.ndst2; ++d)
                if (p.dst2[d] != p.src2) p.dst2[d][i] = v;
        }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(p.done, 1u);
        if (prev == gridDim.x - 1) {
            *p.done = 0;
            __threadfence_system();
            for (int i = 0; i < p.f.n; ++i) st_release_sys(p.f.dst[i], p.value);
        }
    }
}

// computeSNR windows straight from a full spectrum (used when one is available anyway: Parseval variant).
__global__ void spectrum_gather_kernel(const float2* __restrict__ X, int N, const DevResult* __restrict__ res,
                                       float2* __restrict__ sig_win, float2* __restrict__ noise_win) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int len = res->sig_len;
    if (i >= len) return;
    sig_win[i] = X[(res->sig_start + i) & (N - 1)];
    noise_win[i] = X[(res->noise_start + i) & (N - 1)];
}

// ---------------------------------------------------------------------------------------------
// Doppler-rate search dimension: the reference prepares `complexHeterodyne` (kern:755-778,
// dem_base:388) -- out[x] = in[x] * exp(j theta), theta = fmod(((a x) + b) x, 2 pi) + c in fp32 -- and never calls it.
// Same statement here (fp32 phase, the (a x) + b contraction nvcc makes, accurate sinf / cosf), applied to the chunk in the
// time domain so that every rate hypothesis is one more pass of the ordinary Doppler search on the de-chirped chunk.
// ---------------------------------------------------------------------------------------------
__global__ void heterodyne_kernel(const float2* __restrict__ in, float2* __restrict__ out, float a, float b, float c, int n) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    const float xf = (float)x;
    float theta = __fmul_rn(__fmaf_rn(a, xf, b), xf);
    theta = fmodf(theta, 2.0f * 3.14159265358979323846f) + c;
    const float2 w = make_float2(cosf(theta), sinf(theta));
    const float2 v = in[x];
    out[x] = make_float2(__fmaf_rn(v.x, w.x, -__fmul_rn(v.y, w.y)), __fmaf_rn(v.x, w.y, __fmul_rn(v.y, w.x)));   // kern:933-940
}

// fp32 FMA peak probe: 16 independent FMA chains per thread.
__global__ void fma_peak_kernel(float* sink, int iters, float a, float b) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    if (s == 12345.678f) sink[threadIdx.x & 4095] = s;
}

}  // namespace pcs
