/* Plain-C host of the entry points that need no GPU: the bit post-processing (pcs_stitch_*, with the chunk-to-chunk carry
 * handed from one stitcher to another as a multi-rank host does), the computeSNR window mean, the gap filling of the
 * clipped indices and the decoder-side sync search.  Exit code 0 = every check passed.
 *   gcc -std=c99 -I include tests/c_abi/host_only.c -L pycusdr_b200 -lpycusdr_b200 -lm -Wl,-rpath,$PWD/pycusdr_b200 -o host_only */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pycusdr_b200.h"

#define NFFT 4096
#define OVL 1024
#define SPS 16
#define NSYM (NFFT / SPS)

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "check failed: %s (line %d): %s\n", #c, __LINE__, pcs_last_error()); return 1; } } while (0)

static unsigned lcg(unsigned* s) { *s = *s * 1664525u + 1013904223u; return *s >> 8; }

/* symbols of chunk c cut from one bit stream (one symbol per 16 samples, chunks advance by NFFT - OVL samples) */
static int chunk(const uint8_t* stream, int c, int32_t* sym, int32_t* centre, float* mag) {
    const int first = c * (NFFT - OVL) / SPS;
    for (int i = 0; i < NSYM; ++i) { sym[i] = stream[first + i]; centre[i] = i * SPS + 8; mag[i] = 100.0f + i; }
    return NSYM;
}

int main(void) {
    CHECK(pcs_abi_version() == PCS_ABI_VERSION);

    /* --- stitching: one stitcher over five chunks vs. two stitchers taking alternate chunks and passing the carry --- */
    static uint8_t stream[8192];
    unsigned seed = 7;
    for (int i = 0; i < 8192; ++i) stream[i] = (uint8_t)(lcg(&seed) & 1);
    const uint8_t lut[2] = {0, 1};
    pcs_stitch_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.nfft = NFFT; cfg.overlap = OVL; cfg.overlap_offset = 20; cfg.error_threshold = 1000; cfg.match_threshold = 10;
    cfg.num_symbols = 2; cfg.lut_k = 0;
    pcs_stitcher *one = NULL, *even = NULL, *odd = NULL;
    CHECK(pcs_stitch_create(&cfg, lut, NULL, &one) == PCS_OK);
    CHECK(pcs_stitch_create(&cfg, lut, NULL, &even) == PCS_OK);
    CHECK(pcs_stitch_create(&cfg, lut, NULL, &odd) == PCS_OK);
    static int32_t sym[NSYM], centre[NSYM];
    static float mag[NSYM];
    static uint8_t b1[NSYM], c1[NSYM], t1[NSYM], b2[NSYM], c2[NSYM], t2[NSYM], carry[256];
    int32_t n1 = 0, n2 = 0, na = 0, nb = 0, total = 0;
    for (int c = 0; c < 5; ++c) {
        const int n = chunk(stream, c, sym, centre, mag);
        CHECK(pcs_stitch_chunk(one, sym, centre, mag, n, NULL, 0, 16.0, b1, c1, t1, &n1) == PCS_OK);
        pcs_stitcher* mine = (c & 1) ? odd : even;
        pcs_stitcher* other = (c & 1) ? even : odd;
        if (c > 0) {                         /* the carry of chunk c - 1 comes from the stitcher that processed it */
            CHECK(pcs_stitch_get_state(other, carry, sizeof carry, &na, &nb) == PCS_OK);
            CHECK(pcs_stitch_set_state(mine, carry, na, nb) == PCS_OK);
        }
        CHECK(pcs_stitch_chunk(mine, sym, centre, mag, n, NULL, 0, 16.0, b2, c2, t2, &n2) == PCS_OK);
        CHECK(n1 == n2 && n1 > 150);
        CHECK(memcmp(b1, b2, (size_t)n1) == 0 && memcmp(c1, c2, (size_t)n1) == 0 && memcmp(t1, t2, (size_t)n1) == 0);
        /* the window [OVL/2, NFFT - OVL/2] of consecutive chunks tiles the stream: the bits must continue it */
        CHECK(memcmp(b1, stream + OVL / 2 / SPS + (size_t)c * (NFFT - OVL) / SPS, (size_t)n1) == 0);
        total += n1;
    }
    CHECK(total == 5 * (NFFT - OVL) / SPS);
    CHECK(pcs_stitch_get_state(one, carry, 4, &na, &nb) != PCS_OK);          /* buffer too small is an error, not a truncation */
    pcs_stitch_destroy(one); pcs_stitch_destroy(even); pcs_stitch_destroy(odd);

    /* --- computeSNR window mean --- */
    float z[8] = {3, 4, 0, -5, 6, 8, -1, 0}, m = 0;
    CHECK(pcs_mean_abs_c64(z, 4, &m) == PCS_OK && fabsf(m - (5 + 5 + 10 + 1) / 4.0f) < 1e-6f);
    CHECK(pcs_mean_abs_c64(z, 0, &m) != PCS_OK);

    /* --- gap filling of the clipped indices --- */
    const int64_t idx[5] = {10, 11, 14, 200, 299};
    int64_t filled[256];
    int32_t nf = 0;
    CHECK(pcs_fill_gaps(idx, 5, 100, filled, 256, &nf) == PCS_OK);
    CHECK(nf == 5 + 2 + 98 && filled[0] == 10 && filled[2] == 12 && filled[4] == 14 && filled[5] == 200 && filled[nf - 1] == 299);
    CHECK(pcs_fill_gaps(idx, 5, 100, filled, 3, &nf) == PCS_OK && nf == 105);   /* truncated list, full count */

    /* --- decoder-side sync search: a +-1 header found in a bit stream --- */
    const int8_t header[8] = {1, 1, -1, 1, -1, -1, 1, -1};
    int8_t mask[8];
    for (int k = 0; k < 8; ++k) mask[k] = header[7 - k];                      /* protocol.get_mask(): flipped header */
    uint8_t bits[64];
    for (int i = 0; i < 64; ++i) bits[i] = (uint8_t)(lcg(&seed) & 1);
    for (int k = 0; k < 8; ++k) bits[30 + k] = header[k] > 0;
    int32_t at[8], score[8], found = 0;
    CHECK(pcs_sync_search(bits, 64, mask, 8, 4, at, score, 8, &found) == PCS_OK);
    int hit = 0;
    for (int i = 0; i < found && i < 8; ++i) hit |= (at[i] - 8 + 1 == 30 && score[i] == 4);
    CHECK(hit);
    /* plan-time bank factorisation (pcs_factorise_bank + pcs_bank_code_order): four filters = the 2 x 2 sequences of two
     * 32-tap tones; spectra Mk[m][k] = conj(DFT_N(template))[k] by direct summation (N = 256) */
    {
        enum { N = 256, S = 32, M = 4, D = 3 };
        static float masks[M][N][2];
        const double tone[2] = {3.0, -5.0}, pi2 = 6.283185307179586;
        for (int m = 0; m < M; ++m)
            for (int k = 0; k < N; ++k) {
                double re = 0, im = 0;
                for (int n = 0; n < 2 * S; ++n) {
                    const int seg = n / S, t = (m >> seg) & 1;
                    const double ph = pi2 * tone[t] * (n % S) / S + (seg ? 0.7 * (m & 1) : 0.0) - pi2 * k * n / N;
                    re += cos(ph); im += sin(ph);          /* template[n] exp(-2 pi i k n / N) */
                }
                masks[m][k][0] = (float)re; masks[m][k][1] = (float)-im;      /* conj */
            }
        const int32_t shifts[D] = {0, 5, 250};
        int32_t seg = 0, nseg = 0, nbasis = 0, sel[M * PCS_FB_MAX_SEG], form = 0;
        static float coef[2 * D * M * PCS_FB_MAX_SEG], spec[2 * D * PCS_FB_MAX_BASIS * 128];
        /* taps of g = IFFT(Mk) sit at n = 0 and n = -63 .. -1: support_pos 0, support_neg 63 */
        CHECK(pcs_factorise_bank(&masks[0][0][0], N, M, 0, 63, shifts, D, 7, &seg, &nseg, &nbasis, sel, coef, spec) == PCS_OK);
        CHECK(seg == S && nseg == 2 && nbasis == 2);
        CHECK(pcs_bank_code_order(M, nseg, nbasis, D, sel, coef, 1, &form) == PCS_OK);
        CHECK(form == 3);                                   /* complete, and segment 0's constant depends on bit 0 only */
        int seen = 0;
        for (int c = 0; c < M; ++c) seen |= 1 << sel[c];
        CHECK(seen == 15);
        CHECK(pcs_factorise_bank(&masks[0][0][0], N, M, 200, 200, shifts, D, 7, &seg, &nseg, &nbasis, sel, coef, spec) != PCS_OK);
    }
    printf("c abi host-only ok\n");
    return 0;
}
