/* Plain-C host of the drop-in boundary: links libpycusdr_b200.so, builds a tiny configuration (one tone-like matched
 * filter, 8 Doppler bins), pushes one chunk through pcs_upload + pcs_process and prints the result block.
 * Exit codes: 0 ok, 3 no CUDA device (expected on the CPU-only authoring box), anything else = failure.
 *   gcc -std=c99 -I include tests/c_abi/example.c -L pycusdr_b200 -lpycusdr_b200 -lm -Wl,-rpath,$PWD/pycusdr_b200 -o example */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pycusdr_b200.h"

#define N 4096
#define D 8
#define M 2
#define SPS 16
#define TAPS 48

static void dft_conj(const float* re, const float* im, int taps, float* out /* interleaved N */) {
    /* conj(FFT_N(template)) by direct summation (what protocol.get_filter returns, dem_base:196) */
    for (int k = 0; k < N; ++k) {
        double sr = 0, si = 0;
        for (int n = 0; n < taps; ++n) {
            const double a = -2.0 * M_PI * (double)k * n / N;
            sr += re[n] * cos(a) - im[n] * sin(a);
            si += re[n] * sin(a) + im[n] * cos(a);
        }
        out[2 * k] = (float)sr;
        out[2 * k + 1] = (float)-si;
    }
}

int main(void) {
    if (pcs_abi_version() != PCS_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 1; }
    static float masks[M * N * 2], tr[TAPS], ti[TAPS];
    for (int m = 0; m < M; ++m) {           /* two FSK-like templates: +-pi/16 rad per sample */
        for (int n = 0; n < TAPS; ++n) { tr[n] = (float)cos((m ? 1 : -1) * M_PI * n / SPS); ti[n] = (float)sin((m ? 1 : -1) * M_PI * n / SPS); }
        dft_conj(tr, ti, TAPS, masks + (size_t)m * N * 2);
    }
    int32_t shifts[D];
    for (int d = 0; d < D; ++d) shifts[d] = 1000 + 8 * d;
    pcs_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = PCS_ABI_VERSION; cfg.device = 0; cfg.nfft = N; cfg.num_dopplers = D; cfg.num_masks = M;
    cfg.window_width = 7; cfg.sum_all_masks = 1; cfg.samples_per_sym = SPS; cfg.path = PCS_PATH_AUTO; cfg.snr_window = 5;
    pcs_handle* h = NULL;
    int rc = pcs_create(&cfg, shifts, masks, &h);
    if (rc == PCS_ERR_NO_DEVICE) { printf("no CUDA device: %s\n", pcs_last_error()); return 3; }
    if (rc != PCS_OK) { fprintf(stderr, "pcs_create: %s\n", pcs_last_error()); return 1; }
    float* x = (float*)pcs_host_buffer(h);
    unsigned s = 12345u;
    for (int n = 0; n < N; ++n) {           /* alternating-symbol tone pattern at bin shifts[3] plus a little noise */
        const int sym = (n / SPS) & 1;
        const double ph = (sym ? 1 : -1) * M_PI * (n % SPS) / SPS + 2.0 * M_PI * (double)shifts[3] * n / N;
        s = s * 1664525u + 1013904223u;
        const double nz = ((double)(s >> 8) / 16777216.0 - 0.5) * 0.1;
        x[2 * n] = (float)(cos(ph) + nz);
        x[2 * n + 1] = (float)(sin(ph) - nz);
    }
    pcs_result res;
    static float E[D * M], mag[N / (SPS / 2)];
    static int32_t sym[N / (SPS / 2)], centre[N / (SPS / 2)];
    if (pcs_upload(h) != PCS_OK || pcs_process(h, &res, E, sym, centre, mag) != PCS_OK) {
        fprintf(stderr, "process: %s\n", pcs_last_error());
        return 1;
    }
    printf("best_idx %.4f shift %d status %d n_sym %d sp_sym %.4f peak (bin %d mask %d offset %d) E[3][0] %.6g\n",
           res.best_idx, res.shift, res.status, res.n_sym, res.sp_sym, res.peak_bin, res.peak_mask, res.peak_offset, E[3 * M]);
    const int ok = res.status == 0 && res.best_idx >= 0.f && res.best_idx <= (float)(D - 1) && res.n_sym > 200 && fabs(res.sp_sym - SPS) < 1.0;
    if (pcs_demod(h, N, NULL, NULL, NULL, NULL) != PCS_ERR_INVALID) { fprintf(stderr, "bad shift accepted\n"); return 1; }
    pcs_destroy(h);
    if (!ok) { fprintf(stderr, "unexpected result\n"); return 2; }
    printf("c abi example ok\n");
    return 0;
}
