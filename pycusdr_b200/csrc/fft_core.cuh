// In-SM Stockham FFT building blocks (sm_100a, complex fp32).
//
// One *group* of T = B/16 threads transforms B points that live in shared memory; every thread
// owns 16 points in registers per pass, so a B = 16^P transform needs P passes and only P-1
// shared-memory exchanges.  The first pass pulls its inputs through a caller-supplied functor
// (global memory, a fused shift*mask product, ...) and the last pass hands its outputs to a
// caller-supplied sink (shared memory, a fused |y|^2 / arg-max epilogue, ...), which is how the
// Doppler-search kernels keep the correlation surface out of HBM.
//
// Pass structure (autosort, natural order in -> natural order out), current sub-length Ns:
//   j in [0, B/R):  v[r] = in[j + r*B/R];  v[r] *= W_{Ns*R}^{(j mod Ns) * r};  V = DFT_R(v);
//                   out[(j div Ns)*Ns*R + (j mod Ns) + q*Ns] = V[q]
// Shared buffers are indexed through PADI() (one float2 of padding every 16) which makes both
// the strided stores of the early passes and the unit-stride loads conflict-free for 64-bit
// accesses.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcs {

#define PCS_DEVINL __device__ __forceinline__

PCS_DEVINL int padi(int i) { return i + (i >> 4); }
constexpr int padded_len(int n) { return n + (n >> 4); }

// Complex arithmetic on Blackwell's packed fp32x2 pipe forms (FADD2 / FMUL2 / FFMA2, sm_100+): one instruction per
// complex add, subtract or +-i rotation (the swap and per-component negation are operand modifiers in SASS:
// "FADD2 R2, R2.F32x2.HI_LO, -R4.F32x2.LO_HI.NP"), three instead of four per complex multiply.  Every component
// is still an IEEE round-to-nearest add / fma, so results are bit-identical to the scalar formulation; what the
// packed forms buy is issue slots (a packed instruction occupies the FMA pipe for two slots but the scheduler for
// one; measured: tools/ubench/fp32x2.cu).
PCS_DEVINL float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
PCS_DEVINL float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// A complex multiply is two packed instructions, FMUL2 + FFMA2: the broadcast of a.x / a.y ("R.F32"), the swap of b
// ("R.F32x2.LO_HI") and the per-component sign of the addend ("-R.F32x2.HI_LO.NP") are all SASS operand modifiers.
// Keeping every butterfly instruction packed matters beyond the issue slot it saves: a scalar FMUL between packed
// instructions leaves half of the FMA pipe idle for a cycle (tools/ubench/fp32x2.cu: FADD2+FFMA costs 3.5 slots, not 3).
PCS_DEVINL float2 cmul(float2 a, float2 b) {
    // (a.x b.x - a.y b.y, a.x b.y + a.y b.x) = fma2(a.x, b, (-(a.y b.y), a.y b.x)); same roundings as two scalar fmas
    const float2 t = __fmul2_rn(make_float2(a.y, a.y), make_float2(b.y, b.x));
    return __ffma2_rn(make_float2(a.x, a.x), b, make_float2(-t.x, t.y));
}
PCS_DEVINL float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// |z|^2 with the contraction nvcc applies to the reference's ComplexAbsSquared (FMUL + FFMA).
PCS_DEVINL float cabs2(float2 a) { return fmaf(a.x, a.x, a.y * a.y); }

// multiply by exp(i*DIR*theta) given c = cos(theta), s = sin(theta)
template <int DIR>
PCS_DEVINL float2 mulw(float2 v, float c, float s) {
    const float2 t = __fmul2_rn(make_float2(v.y, v.x), make_float2(s, s));
    if (DIR < 0) return __ffma2_rn(v, make_float2(c, c), make_float2(t.x, -t.y));
    return __ffma2_rn(v, make_float2(c, c), make_float2(-t.x, t.y));
}
// multiply by DIR*i
template <int DIR>
PCS_DEVINL float2 muli(float2 v) {
    return DIR < 0 ? make_float2(v.y, -v.x) : make_float2(-v.y, v.x);
}
// a + DIR*i*b and a - DIR*i*b (one FADD2 each)
template <int DIR>
PCS_DEVINL float2 caddi(float2 a, float2 b) { return __fadd2_rn(a, muli<DIR>(b)); }
template <int DIR>
PCS_DEVINL float2 csubi(float2 a, float2 b) { return __fadd2_rn(a, muli<-DIR>(b)); }
// v * (1 + DIR*i) / sqrt(2) and v * (-1 + DIR*i) / sqrt(2)
template <int DIR>
PCS_DEVINL float2 mulw8_1(float2 v) {
    return __fmul2_rn(caddi<DIR>(v, v), make_float2(0.70710678118654752440f, 0.70710678118654752440f));
}
template <int DIR>
PCS_DEVINL float2 mulw8_3(float2 v) {
    return __fmul2_rn(csub(muli<DIR>(v), v), make_float2(0.70710678118654752440f, 0.70710678118654752440f));
}

PCS_DEVINL void dft2(float2& a, float2& b) {
    float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

template <int DIR>
PCS_DEVINL void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = caddi<DIR>(t1, t3);
    a3 = csubi<DIR>(t1, t3);
}

#define PCS_SQRT1_2 0.70710678118654752440f
#define PCS_COS_PI_8 0.92387953251128675613f
#define PCS_SIN_PI_8 0.38268343236508977173f

// In-register DFT of R points.  Output V[q] is left in register slot OutSlot<R>::of(q)... the
// inverse map (slot -> q) is what the callers need and is given by dft_q<R>(slot).
template <int R>
__host__ __device__ constexpr int dft_q(int slot) {
    return R == 16 ? ((slot >> 2) + ((slot & 3) << 2)) : R == 8 ? ((slot >> 2) + ((slot & 3) << 1)) : slot;
}

// slot that holds output q of Dft<R> (inverse of dft_q<R>)
template <int R>
__host__ __device__ constexpr int dft_slot(int q) {
    return R == 16 ? ((q >> 2) + ((q & 3) << 2)) : R == 8 ? ((q >> 1) + ((q & 1) << 2)) : q;
}

template <int R, int DIR>
struct Dft;

template <int DIR>
struct Dft<2, DIR> {
    static PCS_DEVINL void run(float2* v) { dft2(v[0], v[1]); }
};
template <int DIR>
struct Dft<4, DIR> {
    static PCS_DEVINL void run(float2* v) { dft4<DIR>(v[0], v[1], v[2], v[3]); }
};
template <int DIR>
struct Dft<8, DIR> {
    // r = r1*4 + r2 (R1 = 2, R2 = 4); V[q1 + 2*q2] ends in slot 4*q1 + q2
    static PCS_DEVINL void run(float2* v) {
#pragma unroll
        for (int r2 = 0; r2 < 4; ++r2) dft2(v[r2], v[4 + r2]);
        // slot 4 + r2 *= W8^{r2}
        v[5] = mulw8_1<DIR>(v[5]);   // W8^1 = (1 + DIR*i)/sqrt2
        v[6] = muli<DIR>(v[6]);      // W8^2 = DIR*i
        v[7] = mulw8_3<DIR>(v[7]);   // W8^3 = (-1 + DIR*i)/sqrt2
        dft4<DIR>(v[0], v[1], v[2], v[3]);
        dft4<DIR>(v[4], v[5], v[6], v[7]);
    }
};
template <int DIR>
struct Dft<16, DIR> {
    // r = r1*4 + r2 (R1 = R2 = 4); V[q1 + 4*q2] ends in slot 4*q1 + q2
    static PCS_DEVINL void run(float2* v) {
#pragma unroll
        for (int r2 = 0; r2 < 4; ++r2) dft4<DIR>(v[r2], v[4 + r2], v[8 + r2], v[12 + r2]);
        // slot 4*q1 + r2 *= W16^{q1*r2}
        v[5] = mulw<DIR>(v[5], PCS_COS_PI_8, PCS_SIN_PI_8);     // W16^1
        v[6] = mulw8_1<DIR>(v[6]);                              // W16^2
        v[7] = mulw<DIR>(v[7], PCS_SIN_PI_8, PCS_COS_PI_8);     // W16^3
        v[9] = mulw8_1<DIR>(v[9]);                              // W16^2
        v[10] = muli<DIR>(v[10]);                               // W16^4
        v[11] = mulw8_3<DIR>(v[11]);                            // W16^6
        v[13] = mulw<DIR>(v[13], PCS_SIN_PI_8, PCS_COS_PI_8);   // W16^3
        v[14] = mulw8_3<DIR>(v[14]);                            // W16^6
        v[15] = mulw<DIR>(v[15], -PCS_COS_PI_8, -PCS_SIN_PI_8); // W16^9
#pragma unroll
        for (int q1 = 0; q1 < 4; ++q1) dft4<DIR>(v[4 * q1], v[4 * q1 + 1], v[4 * q1 + 2], v[4 * q1 + 3]);
    }
};

// v[r] *= w^r for r = 1..R-1, powers built by multiplication with depth <= 4.
template <int R>
PCS_DEVINL void apply_twiddle_powers(float2* v, float2 w1) {
    if (R == 2) {
        v[1] = cmul(v[1], w1);
        return;
    }
    float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
    v[1] = cmul(v[1], w1);
    v[2] = cmul(v[2], w2);
    v[3] = cmul(v[3], w3);
    if (R >= 8) {
        float2 w4 = cmul(w2, w2);
        v[4] = cmul(v[4], w4);
        v[5] = cmul(v[5], cmul(w4, w1));
        v[6] = cmul(v[6], cmul(w4, w2));
        v[7] = cmul(v[7], cmul(w4, w3));
        if (R >= 16) {
            float2 w8 = cmul(w4, w4);
            v[8] = cmul(v[8], w8);
            v[9] = cmul(v[9], cmul(w8, w1));
            v[10] = cmul(v[10], cmul(w8, w2));
            v[11] = cmul(v[11], cmul(w8, w3));
            float2 w12 = cmul(w8, w4);
            v[12] = cmul(v[12], w12);
            v[13] = cmul(v[13], cmul(w12, w1));
            v[14] = cmul(v[14], cmul(w12, w2));
            v[15] = cmul(v[15], cmul(w12, w3));
        }
    }
}

// exp(-2*pi*i * num / den) for 0 <= num < den <= 2^24 (exact argument reduction in fp32).
PCS_DEVINL float2 unit_phasor_neg(uint32_t num, float inv_den) {
    float s, c;
    sincospif(-2.0f * ((float)num * inv_den), &s, &c);
    return make_float2(c, s);
}

// ---------------------------------------------------------------------------------------------
// Group barrier: T threads of one group (T multiple of 32); id in 1..15 (0 is __syncthreads).
// ---------------------------------------------------------------------------------------------
template <int T>
PCS_DEVINL void group_sync(int bar_id) {
    if (T <= 32) {
        __syncwarp();   // groups of <= 32 threads live inside one warp and run convergent code
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(T) : "memory");
    }
}

template <int LOGB>
struct FftShape {
    static constexpr int B = 1 << LOGB;
    static constexpr int T = B / 16;                    // threads per group
    static constexpr int NPASS = (LOGB + 3) / 4;
    static constexpr int LOG_RLAST = LOGB - 4 * (NPASS - 1);
    static constexpr int RLAST = 1 << LOG_RLAST;
    static constexpr int WORK = padded_len(B);          // float2 per work buffer
};

struct NoPre {
    PCS_DEVINL void operator()(float2*, int) const {}
};

// One pass. R = radix, LOGNS = log2(Ns).  in: idx -> float2.  out: (idx, float2, slot) where slot is
// the compile-time register slot (0..15) of that output inside the thread.  pre(v, j) may modify the
// freshly loaded radix-R vector of butterfly j (used to fuse the Doppler rotation into pass 0).
// Pass twiddles from a table instead of from powers of w1: PassTw<LOGB>::offset(p) + k * R + r holds
// exp(-2 pi i k r / (Ns R)) for pass p (Ns = 16^p), r = 0..R-1, k = 0..Ns-1 (built by the host, build_pass_twiddles).
// One row is R consecutive float2: R/2 128-bit loads that hit L1 replace the log-depth chain of complex multiplies that
// apply_twiddle_powers needs to generate w^2..w^(R-1) (24 of the 87 complex multiplies of a 2048-point transform, and most of
// its fixed-latency stalls).
template <int LOGB>
struct PassTw {
    static constexpr int NPASS = (LOGB + 3) / 4;
    static constexpr int RLAST = 1 << (LOGB - 4 * (NPASS - 1));
    __host__ __device__ static constexpr int radix(int p) { return p == NPASS - 1 ? RLAST : 16; }
    __host__ __device__ static constexpr int offset(int p) {      // float2 elements before pass p's table (pass 0 has none)
        int o = 0;
        for (int q = 1; q < p; ++q) o += (1 << (4 * q)) * radix(q);
        return o;
    }
    __host__ __device__ static constexpr int total() { return offset(NPASS); }
};

template <int LOGB, int R, int LOGNS, int DIR, bool TAB = false, typename In, typename Out, typename Pre = NoPre>
PCS_DEVINL void fft_pass(int t, const float2* __restrict__ tw, In in, Out out, Pre pre = Pre()) {
    constexpr int B = 1 << LOGB, T = B / 16, NB = 16 / R, STR = B / R, NS = 1 << LOGNS;
    constexpr int LOGR = R == 16 ? 4 : R == 8 ? 3 : R == 4 ? 2 : 1;
    float2 v[NB][R];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int j = t + b * T;
#pragma unroll
        for (int r = 0; r < R; ++r) v[b][r] = in(j + r * STR);
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int j = t + b * T;
        const int k = j & (NS - 1);
        pre(v[b], j);
        if (LOGNS > 0) {
            if (TAB) {                                      // tw = this pass's table
                const float4* __restrict__ row = reinterpret_cast<const float4*>(tw + (size_t)k * R);
#pragma unroll
                for (int rr = 0; rr < R / 2; ++rr) {
                    const float4 w = __ldg(&row[rr]);
                    if (rr > 0) v[b][2 * rr] = DIR < 0 ? cmul(v[b][2 * rr], make_float2(w.x, w.y))
                                                       : cmul(v[b][2 * rr], make_float2(w.x, -w.y));
                    v[b][2 * rr + 1] = DIR < 0 ? cmul(v[b][2 * rr + 1], make_float2(w.z, w.w))
                                               : cmul(v[b][2 * rr + 1], make_float2(w.z, -w.w));
                }
            } else {
                float2 w1 = __ldg(&tw[k * (B / (NS * R))]);   // forward table exp(-2 pi i t / B)
                if (DIR > 0) w1 = cconj(w1);
                apply_twiddle_powers<R>(v[b], w1);
            }
        }
        Dft<R, DIR>::run(v[b]);
        const int j0 = ((j >> LOGNS) << (LOGNS + LOGR)) + k;
#pragma unroll
        for (int s = 0; s < R; ++s) out(j0 + dft_q<R>(s) * NS, v[b][s], b * R + s);
    }
}

// Whole transform for one group.  ``work0``/``work1`` are the group's private shared buffers of
// FftShape<LOGB>::WORK float2 each.  Pass p stores to work[p & 1] and pass p+1 loads from it, so
// NPASS-1 barriers are enough inside one transform; back-to-back transforms on the same buffers
// are safe for odd NPASS, and one trailing barrier is added for even NPASS (the last pass then
// still reads work0 while the next transform's pass 0 would overwrite it).
// src(idx) supplies natural-order input idx, sink(idx, val, slot) receives natural-order output idx.
template <int LOGB, int DIR, bool TAB = false, typename Src, typename Sink, typename Pre = NoPre>
PCS_DEVINL void group_fft(float2* work0, float2* work1, const float2* __restrict__ tw, int t, int bar_id,
                          Src src, Sink sink, Pre pre = Pre()) {
    using S = FftShape<LOGB>;
    using P = PassTw<LOGB>;
    auto ld0 = [&](int i) { return work0[padi(i)]; };
    auto ld1 = [&](int i) { return work1[padi(i)]; };
    auto st0 = [&](int i, float2 v, int) { work0[padi(i)] = v; };
    auto st1 = [&](int i, float2 v, int) { work1[padi(i)] = v; };
    static_assert(S::NPASS >= 2 && S::NPASS <= 4, "supported transform sizes: 2^5 .. 2^16");
    // TAB: tw points at the pass tables (PassTw<LOGB>), else at the forward table exp(-2 pi i t / B)
    constexpr int O1 = P::offset(1), O2 = P::offset(2), O3 = P::offset(3);
    const float2* tw1 = TAB ? tw + O1 : tw;
    const float2* tw2 = TAB ? tw + O2 : tw;
    const float2* tw3 = TAB ? tw + O3 : tw;
    if constexpr (S::NPASS == 2) {
        fft_pass<LOGB, 16, 0, DIR, false>(t, tw, src, st0, pre);
        group_sync<S::T>(bar_id);
        fft_pass<LOGB, S::RLAST, 4, DIR, TAB>(t, tw1, ld0, sink);
        group_sync<S::T>(bar_id);
    } else if constexpr (S::NPASS == 3) {
        fft_pass<LOGB, 16, 0, DIR, false>(t, tw, src, st0, pre);
        group_sync<S::T>(bar_id);
        fft_pass<LOGB, 16, 4, DIR, TAB>(t, tw1, ld0, st1);
        group_sync<S::T>(bar_id);
        fft_pass<LOGB, S::RLAST, 8, DIR, TAB>(t, tw2, ld1, sink);
    } else {
        fft_pass<LOGB, 16, 0, DIR, false>(t, tw, src, st0, pre);
        group_sync<S::T>(bar_id);
        fft_pass<LOGB, 16, 4, DIR, TAB>(t, tw1, ld0, st1);
        group_sync<S::T>(bar_id);
        fft_pass<LOGB, 16, 8, DIR, TAB>(t, tw2, ld1, st0);
        group_sync<S::T>(bar_id);
        fft_pass<LOGB, S::RLAST, 12, DIR, TAB>(t, tw3, ld0, sink);
        group_sync<S::T>(bar_id);
    }
}

}  // namespace pcs
