// Segment factorisation of a matched-filter bank (plan time, host only, no CUDA call).
//
// The search correlates every Doppler-shifted block with M filters of L taps (a6-a8: kern:339-373, 421-480).  Banks built
// from piecewise templates -- the FSK-2 bank of the reference is M = 2^k concatenations of k one-symbol tone segments with a
// continuous phase (pyCuSDR/protocol/FSK2_base.py:17-46: 8 masks x 3 symbols x 128 samples for CC11xx) -- have far fewer
// DISTINCT segments than M * k: every segment is a complex constant times one of R basis segments (R = 2 tones for FSK-2).
// With the taps g_m[n] (n = -Lneg .. Lpos) cut into J segments of S = L / J taps counted from the Lpos end,
//
//     g_m[n] = sum_j c[m][j] * b_{sel[m][j]}[n + j S],          b_r supported on n = Lpos - S + 1 .. Lpos,
//
// the M filter outputs are J-term combinations of R filter outputs taken j S samples later:
//
//     y_m[i] = (g_m * x)[i] = sum_j c[m][j] * u_{sel[m][j]}[i + j S],        u_r = b_r * x,
//
// so a (bin, block) item needs R inverse transforms instead of M (search_fb_kernel).  A Doppler shift s_d multiplies the
// taps by exp(2 pi i s_d n / N): the basis spectra become per-bin tables and c picks up exp(-2 pi i s_d j S / N).
//
// Nothing here is specific to FSK: the structure is DETECTED from the spectra pcs_create is given (greedy assignment of the
// M * J segments to basis vectors, least-squares refit, residual test), and the factorised bank is accepted only if it
// reproduces the per-bin filter spectra the unfactorised kernel would use to 1e-5 of their peak.
#include <math.h>
#include <stdint.h>

#include <complex>
#include <vector>

#include "../../include/pycusdr_b200.h"

int pcs_fail_msg(int code, const char* msg);

namespace {
typedef std::complex<double> cd;

// Iterative radix-2 transform in double precision; dir = -1 forward, +1 inverse (unnormalised).
void fft_pow2(std::vector<cd>& a, int dir) {
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    std::vector<cd> w(n / 2);
    for (size_t k = 0; k < n / 2; ++k) {
        const double ang = dir * 2.0 * M_PI * (double)k / (double)n;
        w[k] = cd(cos(ang), sin(ang));
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const size_t half = len >> 1, step = n / len;
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < half; ++k) {
                const cd u = a[i + k], v = a[i + k + half] * w[k * step];
                a[i + k] = u + v;
                a[i + k + half] = u - v;
            }
    }
}

cd unit(long long num, long long den) {      // exp(2 pi i num / den), argument reduced exactly
    long long r = num % den;
    if (r < 0) r += den;
    const double ang = 2.0 * M_PI * (double)r / (double)den;
    return cd(cos(ang), sin(ang));
}

struct Factor {
    int S = 0, J = 0, R = 0;
    std::vector<int> sel;              // [M][J]
    std::vector<cd> c;                 // [M][J]
    std::vector<std::vector<cd>> b;    // R x S
};

// taps[m][a], a = 0 .. L-1 <-> n = a - Lneg.  Segment j = taps[L - (j+1) S .. L - j S).
bool factorise(const std::vector<std::vector<cd>>& taps, int L, int J, int max_basis, double tol, Factor* out) {
    const int M = (int)taps.size();
    if (L % J) return false;
    const int S = L / J;
    Factor f;
    f.S = S; f.J = J;
    f.sel.assign((size_t)M * J, 0);
    f.c.assign((size_t)M * J, cd(0, 0));
    double max_norm = 0;
    for (int m = 0; m < M; ++m)
        for (int a = 0; a < L; ++a) max_norm = std::max(max_norm, std::norm(taps[m][a]));
    max_norm *= S;
    for (int m = 0; m < M; ++m)
        for (int j = 0; j < J; ++j) {
            const cd* v = &taps[m][(size_t)L - (size_t)(j + 1) * S];
            double vn = 0;
            for (int i = 0; i < S; ++i) vn += std::norm(v[i]);
            if (vn <= 1e-14 * max_norm) continue;        // an empty segment: coefficient 0
            int found = -1;
            cd cf(0, 0);
            for (int r = 0; r < (int)f.b.size() && found < 0; ++r) {
                cd ip(0, 0);
                for (int i = 0; i < S; ++i) ip += std::conj(f.b[r][i]) * v[i];      // basis vectors have unit norm
                double res = 0;
                for (int i = 0; i < S; ++i) res += std::norm(v[i] - ip * f.b[r][i]);
                if (res <= tol * tol * vn) { found = r; cf = ip; }
            }
            if (found < 0) {
                if ((int)f.b.size() >= max_basis) return false;
                const double inv = 1.0 / sqrt(vn);
                std::vector<cd> nb((size_t)S);
                for (int i = 0; i < S; ++i) nb[i] = v[i] * inv;
                f.b.push_back(nb);
                found = (int)f.b.size() - 1;
                cf = cd(sqrt(vn), 0);
            }
            f.sel[(size_t)m * J + j] = found;
            f.c[(size_t)m * J + j] = cf;
        }
    f.R = (int)f.b.size();
    if (f.R == 0) return false;
    // least-squares refit of every basis vector over the segments assigned to it, then of the coefficients
    for (int r = 0; r < f.R; ++r) {
        std::vector<cd> acc((size_t)S, cd(0, 0));
        double den = 0;
        for (int m = 0; m < M; ++m)
            for (int j = 0; j < J; ++j) {
                const size_t q = (size_t)m * J + j;
                if (f.sel[q] != r || f.c[q] == cd(0, 0)) continue;
                const cd* v = &taps[m][(size_t)L - (size_t)(j + 1) * S];
                for (int i = 0; i < S; ++i) acc[i] += std::conj(f.c[q]) * v[i];
                den += std::norm(f.c[q]);
            }
        if (den <= 0) continue;
        double bn = 0;
        for (int i = 0; i < S; ++i) { acc[i] /= den; bn += std::norm(acc[i]); }
        const double inv = 1.0 / sqrt(bn);
        for (int i = 0; i < S; ++i) f.b[r][i] = acc[i] * inv;
    }
    for (int m = 0; m < M; ++m) {
        double gn = 0, res = 0;
        for (int j = 0; j < J; ++j) {
            const size_t q = (size_t)m * J + j;
            const cd* v = &taps[m][(size_t)L - (size_t)(j + 1) * S];
            const std::vector<cd>& b = f.b[f.sel[q]];
            cd ip(0, 0);
            for (int i = 0; i < S; ++i) ip += std::conj(b[i]) * v[i];
            if (f.c[q] != cd(0, 0)) f.c[q] = ip;
            for (int i = 0; i < S; ++i) {
                gn += std::norm(v[i]);
                res += std::norm(v[i] - f.c[q] * b[i]);
            }
        }
        if (res > tol * tol * gn) return false;
    }
    *out = f;
    return true;
}
}  // namespace

// See include/pycusdr_b200.h.  *num_basis = 0 on return means "keep the unfactorised bank".
int pcs_factorise_bank(const float* masks, int32_t nfft, int32_t num_masks, int32_t support_pos, int32_t support_neg,
                       const int32_t* shifts, int32_t num_shifts, int32_t log2_block, int32_t* seg_len, int32_t* num_seg,
                       int32_t* num_basis, int32_t* sel_out, float* coef_out, float* basis_spec_out) {
    if (!masks || !shifts || !seg_len || !num_seg || !num_basis) return pcs_fail_msg(PCS_ERR_INVALID, "null argument");
    const int N = nfft, M = num_masks, D = num_shifts, Lpos = support_pos, Lneg = support_neg, L = Lpos + Lneg + 1;
    *seg_len = *num_seg = *num_basis = 0;
    if (N < 2 || (N & (N - 1)) || M < 2 || D < 1 || Lpos < 0 || Lneg < 0 || log2_block < 5 || (1 << log2_block) > N || L > (1 << log2_block))
        return pcs_fail_msg(PCS_ERR_INVALID, "pcs_factorise_bank: bad geometry");
    const int B = 1 << log2_block;
    const std::complex<float>* mk = reinterpret_cast<const std::complex<float>*>(masks);

    // taps in the units of an unnormalised inverse transform: g_un[n] = sum_k Mk[k] exp(2 pi i k n / N)
    std::vector<std::vector<cd>> taps((size_t)M, std::vector<cd>((size_t)L));
    {
        std::vector<cd> a((size_t)N);
        for (int m = 0; m < M; ++m) {
            for (int k = 0; k < N; ++k) a[k] = cd(mk[(size_t)m * N + k]);
            fft_pow2(a, +1);
            for (int i = 0; i < L; ++i) taps[m][i] = a[(size_t)((i - Lneg) & (N - 1))];
        }
    }
    // fewest transforms + combination terms; a transform of the generic kernel costs about 22 combination terms
    // (1566 against ~70 SM cycles per (bin, block) item on C1: DESIGN.md section 3)
    Factor best;
    double best_cost = 0.75 * M;
    for (int J = 2; J <= PCS_FB_MAX_SEG; ++J) {
        Factor f;
        if (L % J || ((L / J) & 1) || L / J < Lpos + 2) continue;       // S even (128-bit loads of output pairs), S >= Lpos + 2
        if (!factorise(taps, L, J, PCS_FB_MAX_BASIS, 2e-6, &f)) continue;
        const double cost = f.R + 0.045 * M * J;
        if (cost < best_cost) { best_cost = cost; best = f; }
    }
    if (best.R == 0) return PCS_OK;
    const int S = best.S, J = best.J, R = best.R;

    // per-bin tables
    std::vector<cd> spec((size_t)D * R * B), coef((size_t)D * M * J);
    std::vector<cd> w((size_t)B);
    for (int d = 0; d < D; ++d) {
        const long long s = shifts[d];
        for (int r = 0; r < R; ++r) {
            std::fill(w.begin(), w.end(), cd(0, 0));
            for (int i = 0; i < S; ++i) {
                const int n = Lpos - S + 1 + i;
                w[(size_t)(n & (B - 1))] = best.b[r][i] * unit(s * n, N);
            }
            fft_pow2(w, -1);
            for (int k = 0; k < B; ++k) spec[((size_t)d * R + r) * B + k] = w[k] / (double)B;
        }
        for (int m = 0; m < M; ++m)
            for (int j = 0; j < J; ++j) coef[((size_t)d * M + m) * J + j] = best.c[(size_t)m * J + j] * unit(-s * j * S, N);
    }
    // acceptance test: the factorised bank must reproduce G[d][m][k] = Mk[m][(k N/B - s_d) % N] * N/B
    {
        const int dec = N / B;
        const int probe[3] = {0, D / 2, D - 1};
        double err = 0, peak = 0;
        for (int pi = 0; pi < 3; ++pi) {
            const int d = probe[pi];
            for (int m = 0; m < M; ++m)
                for (int k = 0; k < B; ++k) {
                    cd g(0, 0);
                    for (int j = 0; j < J; ++j)
                        g += coef[((size_t)d * M + m) * J + j] * spec[((size_t)d * R + best.sel[(size_t)m * J + j]) * B + k] *
                             unit((long long)k * j * S, B);
                    const cd ref = cd(mk[(size_t)m * N + (size_t)(((long long)k * dec - shifts[d]) & (N - 1))]) * (double)dec;
                    err = std::max(err, std::abs(g - ref));
                    peak = std::max(peak, std::abs(ref));
                }
        }
        if (!(err <= 1e-5 * peak)) return PCS_OK;
    }
    *seg_len = S; *num_seg = J; *num_basis = R;
    if (sel_out)
        for (size_t q = 0; q < (size_t)M * J; ++q) sel_out[q] = best.sel[q];
    if (coef_out)
        for (size_t q = 0; q < coef.size(); ++q) { coef_out[2 * q] = (float)coef[q].real(); coef_out[2 * q + 1] = (float)coef[q].imag(); }
    if (basis_spec_out)
        for (size_t q = 0; q < spec.size(); ++q) { basis_spec_out[2 * q] = (float)spec[q].real(); basis_spec_out[2 * q + 1] = (float)spec[q].imag(); }
    return PCS_OK;
}

// See include/pycusdr_b200.h.  Decides which form of search_fb_kernel a factorised bank can take and, for a complete binary
// bank, puts the tables into the kernel's order.
int pcs_bank_code_order(int32_t num_masks, int32_t num_seg, int32_t num_basis, int32_t num_shifts, int32_t* sel, float* coef,
                        int32_t allow_shared_sums, int32_t* form) {
    if (!sel || !coef || !form || num_masks < 1 || num_seg < 1 || num_seg > PCS_FB_MAX_SEG || num_basis < 1 || num_shifts < 1)
        return pcs_fail_msg(PCS_ERR_INVALID, "pcs_bank_code_order: bad argument");
    const int M = num_masks, J = num_seg, D = num_shifts;
    *form = 1;
    if (num_basis != 2 || J > 3 || M != (1 << J)) return PCS_OK;
    std::vector<int32_t> code_mask((size_t)M, -1);
    for (int m = 0; m < M; ++m) {
        int code = 0;
        for (int j = 0; j < J; ++j) {
            const int r = sel[(size_t)m * J + j];
            if (r < 0 || r > 1) return pcs_fail_msg(PCS_ERR_INVALID, "pcs_bank_code_order: selector outside the basis");
            code |= r << j;
        }
        if (code_mask[code] >= 0) return PCS_OK;          // two filters with the same selectors: not complete
        code_mask[code] = m;
    }
    std::vector<float> cc((size_t)2 * D * M * J);
    for (int d = 0; d < D; ++d)
        for (int code = 0; code < M; ++code)
            for (int q = 0; q < 2 * J; ++q)
                cc[2 * (((size_t)d * M + code) * J) + q] = coef[2 * (((size_t)d * M + code_mask[code]) * J) + q];
    for (size_t q = 0; q < cc.size(); ++q) coef[q] = cc[q];
    for (int code = 0; code < M; ++code) sel[code] = code_mask[code];
    *form = 2;
    if (!allow_shared_sums) return PCS_OK;
    for (size_t d = 0; d < (size_t)D; ++d)
        for (int code = 0; code < M; ++code)
            for (int j = 0; j + 1 < J; ++j) {
                const float* a = &coef[2 * ((d * M + code) * J + j)];
                const float* b = &coef[2 * ((d * M + (code & ((2 << j) - 1))) * J + j)];
                const float tol = 1e-6f * (fabsf(b[0]) + fabsf(b[1]));
                if (fabsf(a[0] - b[0]) > tol || fabsf(a[1] - b[1]) > tol) return PCS_OK;
            }
    *form = 3;
    return PCS_OK;
}
