// C ABI of the demodulator hot path: handle management, planning and kernel orchestration.
// Declarations and the reference interfaces they replace: include/pycusdr_b200.h.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/pycusdr_b200.h"
#include "kernels.cuh"

using namespace pcs;

#define PCS_MAX_DEVICES 64
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

int pcs_fail_msg(int code, const char* msg) { return fail(code, "%s", msg); }   // for the other translation units

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fail(PCS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                   \
    } while (0)

// Functors for the tiled FFT passes ------------------------------------------------------------
struct LoadC {
    const float2* p;
    __device__ __forceinline__ float2 operator()(long long i) const { return __ldg(&p[i]); }
};
struct LoadR {
    const float* p;
    __device__ __forceinline__ float2 operator()(long long i) const { return make_float2(__ldg(&p[i]), 0.f); }
};
struct StoreC {
    float2* p;
    __device__ __forceinline__ void operator()(long long i, float2 v) const { p[i] = v; }
};

struct pcs_handle {
    pcs_config cfg;
    cudaStream_t stream = nullptr;
    int N = 0, logN = 0, D = 0, M = 0, sm_count = 0;
    int max_sym = 0, spsym_min = 0, i_high = 0, i_low = 0;
    // overlap-save plan
    int path = 0, logB = 0, G = 1, nblk = 0, V = 0, Lpos = 0, Lneg = 0, search_ctas = 0, search_smem = 0;
    // large-FFT plan
    int logN1 = 0, logN2 = 0;
    std::map<int, float2*> tw;     // forward twiddle tables by log2 size
    std::map<int, float2*> twp;    // per-pass twiddle tables (PassTw<LOGB>) by log2 size
    // working set of one in-flight generic (shared-memory) search: partials, winning blocks, block spectra; a second set lets
    // two chunks' searches overlap on two streams (pcs_shard_*), cur_lane selects the one the next enqueue uses
    struct OsBufs {
        float *psum = nullptr, *pmax = nullptr;   // [D][M][nblk]
        int* wblk = nullptr;                      // [D][M] block of the largest |y|^2
        float2* xbs = nullptr;                    // [nblk][B] block spectra (shifted-filter form)
    } osb[2];
    // device memory
    std::vector<void*> dev_allocs;
    int64_t dev_bytes = 0;
    float2 *d_x = nullptr, *d_X = nullptr, *d_masks = nullptr, *d_gb = nullptr, *d_scratch = nullptr;
    float2 *d_Pf = nullptr, *d_ycplx = nullptr, *d_sigwin = nullptr, *d_noisewin = nullptr;
    const float2* d_x_cur = nullptr;   // chunk source of the current upload (d_x or external)
    const float2* d_x_base = nullptr;  // the chunk as uploaded (pcs_heterodyne de-chirps it into d_xh and points d_x_cur there)
    float2* d_xh = nullptr;
    int* d_shifts = nullptr;
    std::vector<int32_t> h_shifts;     // host copy of the shift table (computeSNR window geometry)
    float *d_Efull = nullptr, *d_E = nullptr, *d_peakv = nullptr;
    int* d_peako = nullptr;
    float *d_ymag = nullptr, *d_p = nullptr, *d_mag = nullptr;
    int *d_sym = nullptr, *d_centre = nullptr;
    DevResult* d_res = nullptr;
    // d_res | d_E | d_sym | d_centre | d_mag | d_sigwin | d_noisewin live in ONE allocation so that a chunk's results can
    // leave with a single D2H copy (pcs_shard_*); rb_off[] = byte offsets of the seven parts, rb_bytes = total
    unsigned char* d_rblock = nullptr;
    size_t rb_off[7] = {}, rb_bytes = 0;
    // pinned host memory
    float2 *h_x = nullptr, *h_sigwin = nullptr, *h_noisewin = nullptr;
    const float2* h_src = nullptr;     // one-shot: the next pcs_upload copies from this page-locked caller buffer instead of h_x
    DevResult* h_res = nullptr;
    float *h_E = nullptr, *h_mag = nullptr;
    int *h_sym = nullptr, *h_centre = nullptr;
    bool uploaded = false, searched = false, demodulated = false;
    bool own_stream = true;
    // register-resident B = 256 search plan (short filters)
    bool fast256 = false;
    int nblk256 = 0, V256 = 0;
    float4* d_gperm = nullptr;
    float *d_thr_partial = nullptr, *d_thr_level = nullptr;   // pcs_upload_thresholded (allocated on first use)
    unsigned int *d_thr_bits = nullptr, *h_thr_bits = nullptr;
    bool os_fs = false;                // shifted-filter form of the generic search
    bool fs256 = false;                // shifted-filter form of the 256-point search (block spectra shared by all bins)
    float4* d_gs = nullptr;
    // per-launch working set of the shifted-filter search; a second set lets two chunks' searches overlap on two streams
    // (bin sharding: pcs_shard_*), cur_lane selects the one the next enqueue uses
    struct Fs256Bufs {
        float4* xbs = nullptr;                       // block spectra of the chunk
        float *psum = nullptr, *pmax = nullptr;      // [D][nblk][M] per-item partials
        unsigned int *bin_count = nullptr, *bins_done = nullptr;
    } fsb[2];
    int cur_lane = 0;
    const float4* xbs_ext = nullptr;   // when set, the next shifted-filter search reads these block spectra instead of computing them
    float2* d_gs_os = nullptr;         // per-bin filter spectra of the generic kernel's shifted-filter form (natural order)
    // factorised filter bank (bank_factor.cu): R basis filters, J segments of S taps per filter; search_fb_kernel
    bool fb = false;
    int fb_complete = 0;     // 1: R = 2 and the M = 2^J selector rows are all different (an FSK-2 bank); 2: and the coefficient of
                             // segment j depends on the selectors of segments 0..j only (shared partial sums)
    int fb_S = 0, fb_J = 0, fb_R = 0;
    float2 *d_fb_basis = nullptr, *d_fb_coef = nullptr;   // [D][R][B], [D][M][J]
    int* d_fb_sel = nullptr;                               // [M][J]
    int fs_items = 0;                  // items (bin, block) per CTA; 0 = choose per launch
    int fs_items_cap = 0;              // upper bound for the chosen value (0 = none)
    float *d_psum256 = nullptr, *d_pmax256 = nullptr, *d_part_sum = nullptr, *d_part_max = nullptr;   // rotate-form kernels
    int* d_part_blk = nullptr;
    float2* d_scratch2 = nullptr;      // pass-1 output of the timing-recovery transform
    // Parseval variant (labelled alternative: energies only, no peak)
    bool parseval = false;
    int pv_mw = 0;                     // weight rows: 1 in SUM mode, else M
    float *d_W = nullptr, *d_PX = nullptr, *d_pvpart = nullptr;
    bool spectrum_full = false;        // d_X holds the full spectrum of the current chunk (computed on demand)
    // CUDA graph of the per-chunk sequence search -> estimate -> demod -> result copies (captured after one eager run)
    cudaGraphExec_t gexec = nullptr;
    bool graph_enabled = true, graph_failed = false, fetch_in_flight = false;
    int eager_chunks = 0;
    int64_t graph_launches = 0;
    // bin-sharded streaming search over NVLink peer memory (one process per GPU, pcs_shard_*): see the section at the
    // end of this file for the layout of the exchange region and the flag protocol
    struct Shard* shard = nullptr;
    unsigned int* d_done = nullptr;                       // CTA completion counter of the locate kernel (rotate-form search)
    unsigned long long* push_flag = nullptr;              // when set, the search stage publishes its arrival there
    unsigned long long push_value = 0;
    unsigned long long* push_ack = nullptr;               // optional second flag raised with push_flag (fused search only)
    float *tab_E = nullptr, *tab_pv = nullptr;   // tables the search stage writes / the estimate stage reads
    int* tab_po = nullptr;
    cudaStream_t side = nullptr;       // forked branch of the graph: chunk spectrum -> SNR bins (off the critical path)
    cudaEvent_t ev_fork = nullptr, ev_est = nullptr, ev_side = nullptr;
    bool spectrum_pending = false;     // pass 1 of the chunk spectrum not enqueued yet for the current upload
    int bin_lo = 0, bin_hi = 0;        // Doppler rows this handle searches (bin sharding); default all
    int win_cap = PCS_WINDOW_MAX;      // largest computeSNR window this Doppler grid can produce
    int64_t launches = 0;
    // optional per-stage CUDA-event timing (pcs_set_profiling)
    bool profiling = false;
    cudaEvent_t ev[2 * PCS_NUM_STAGES] = {};
    bool ev_pending[PCS_NUM_STAGES] = {};
    double stage_ms[PCS_NUM_STAGES] = {};
    int64_t stage_count[PCS_NUM_STAGES] = {};
};

// Stage timer: records a CUDA event pair on the handle's stream around a stage when profiling is on.
struct StageTimer {
    pcs_handle* h;
    int stage;
    StageTimer(pcs_handle* h_, int stage_) : h(h_), stage(stage_) {
        if (!h->profiling) return;
        if (h->ev_pending[stage]) {   // harvest the previous measurement of this stage
            float ms = 0.f;
            if (cudaEventSynchronize(h->ev[2 * stage + 1]) == cudaSuccess &&
                cudaEventElapsedTime(&ms, h->ev[2 * stage], h->ev[2 * stage + 1]) == cudaSuccess) {
                h->stage_ms[stage] += ms;
                h->stage_count[stage]++;
            }
            h->ev_pending[stage] = false;
        }
        cudaEventRecord(h->ev[2 * stage], h->stream);
    }
    ~StageTimer() {
        if (!h->profiling) return;
        cudaEventRecord(h->ev[2 * stage + 1], h->stream);
        h->ev_pending[stage] = true;
    }
};

template <typename T>
static int dev_alloc(pcs_handle* h, T** p, size_t count) {
    void* q = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CUDA_TRY(cudaMalloc(&q, bytes));
    h->dev_allocs.push_back(q);
    h->dev_bytes += (int64_t)bytes;
    *p = reinterpret_cast<T*>(q);
    return 0;
}

static int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

static int get_twiddles(pcs_handle* h, int logB, const float2** out) {
    auto it = h->tw.find(logB);
    if (it == h->tw.end()) {
        const int B = 1 << logB;
        std::vector<float2> host(B);
        for (int t = 0; t < B; ++t) {
            const double a = -2.0 * M_PI * (double)t / (double)B;
            host[t] = make_float2((float)cos(a), (float)sin(a));
        }
        float2* d = nullptr;
        if (int rc = dev_alloc(h, &d, (size_t)B)) return rc;
        CUDA_TRY(cudaMemcpyAsync(d, host.data(), sizeof(float2) * B, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        it = h->tw.emplace(logB, d).first;
    }
    *out = it->second;
    return 0;
}

// Pass tables of the shared-memory transforms (fft_core.cuh: PassTw): for pass p (Ns = 16^p, radix R) the row k holds
// exp(-2 pi i k r / (Ns R)), r = 0..R-1, in float32 from float64.
template <int LOGB>
static void fill_pass_twiddles(std::vector<float2>& host) {
    using P = PassTw<LOGB>;
    host.assign((size_t)P::total(), make_float2(1.f, 0.f));
    for (int p = 1; p < P::NPASS; ++p) {
        const int Ns = 1 << (4 * p), R = P::radix(p);
        for (int k = 0; k < Ns; ++k)
            for (int r = 0; r < R; ++r) {
                const double a = -2.0 * M_PI * (double)k * (double)r / ((double)Ns * (double)R);
                host[(size_t)P::offset(p) + (size_t)k * R + r] = make_float2((float)cos(a), (float)sin(a));
            }
    }
}

static int get_pass_twiddles(pcs_handle* h, int logB, const float2** out) {
    auto it = h->twp.find(logB);
    if (it == h->twp.end()) {
        std::vector<float2> host;
        switch (logB) {
            case 9: fill_pass_twiddles<9>(host); break;
            case 10: fill_pass_twiddles<10>(host); break;
            case 11: fill_pass_twiddles<11>(host); break;
            case 12: fill_pass_twiddles<12>(host); break;
            case 13: fill_pass_twiddles<13>(host); break;
            default: return fail(PCS_ERR_INVALID, "no pass twiddles for 2^%d", logB);
        }
        float2* d = nullptr;
        if (int rc = dev_alloc(h, &d, host.size())) return rc;
        CUDA_TRY(cudaMemcpyAsync(d, host.data(), sizeof(float2) * host.size(), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        it = h->twp.emplace(logB, d).first;
    }
    *out = it->second;
    return 0;
}

// ---- tiled FFT pass launcher ------------------------------------------------------------------
template <int LOGB, int DIR, int C, typename Load, typename Store>
static int launch_tile_t(pcs_handle* h, const TileGeom& g, Load ld, Store st) {
    using S = FftShape<LOGB>;
    static_assert(S::T <= 32 || C <= 15, "named barrier ids");
    const float2* tw = nullptr;
    if (int rc = get_twiddles(h, LOGB, &tw)) return rc;
    const size_t smem = (size_t)2 * C * S::WORK * sizeof(float2);
    auto kern = fft_tile_kernel<LOGB, DIR, C, Load, Store>;
    static bool configured[PCS_MAX_DEVICES] = {};   // per instantiation and device (the attribute is per device)
    if (!configured[h->cfg.device]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[h->cfg.device] = true;
    }
    const int grid = (g.nvec + C - 1) / C;
    kern<<<grid, C * S::T, smem, h->stream>>>(g, tw, ld, st);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

template <int DIR, typename Load, typename Store>
static int launch_tile(pcs_handle* h, int logB, const TileGeom& g, Load ld, Store st) {
    switch (logB) {
        case 6: return launch_tile_t<6, DIR, 32>(h, g, ld, st);
        case 7: return launch_tile_t<7, DIR, 32>(h, g, ld, st);
        case 8: return launch_tile_t<8, DIR, 16>(h, g, ld, st);
        case 9: return launch_tile_t<9, DIR, 16>(h, g, ld, st);
        case 10: return launch_tile_t<10, DIR, 8>(h, g, ld, st);
        case 11: return launch_tile_t<11, DIR, 4>(h, g, ld, st);
        case 12: return launch_tile_t<12, DIR, 2>(h, g, ld, st);
    }
    return fail(PCS_ERR_INVALID, "unsupported tile transform size 2^%d", logB);
}

// Same transform sizes with 4 vectors per CTA: four times the CTAs when a pass has too few vectors to fill the GPU.
template <int DIR, typename Load, typename Store>
static int launch_tile_narrow(pcs_handle* h, int logB, const TileGeom& g, Load ld, Store st) {
    switch (logB) {
        case 6: return launch_tile_t<6, DIR, 4>(h, g, ld, st);
        case 7: return launch_tile_t<7, DIR, 4>(h, g, ld, st);
        case 8: return launch_tile_t<8, DIR, 4>(h, g, ld, st);
        case 9: return launch_tile_t<9, DIR, 4>(h, g, ld, st);
    }
    return launch_tile<DIR>(h, logB, g, ld, st);
}

// First pass of the four-step N-point transform: S[k1][n2] = W_N^(n2 k1) * sum_n1 in[n1 N2 + n2] W_N1^(n1 k1).
template <int DIR, typename Load>
static int fft_pass1(pcs_handle* h, Load ld, float2* S) {
    const int N2 = 1 << h->logN2;
    TileGeom g1{};
    g1.nvec = N2;
    g1.in_vs = 1; g1.in_es = N2; g1.out_vs = 1; g1.out_es = N2;
    g1.v_fast_in = 1; g1.v_fast_out = 1;
    g1.twiddle = 1; g1.inv_ntw = 1.0f / (float)h->N; g1.ntw_mask = (uint32_t)h->N - 1u;
    if (N2 / 16 < h->sm_count / 2) return launch_tile_narrow<DIR>(h, h->logN1, g1, ld, StoreC{S});
    return launch_tile<DIR>(h, h->logN1, g1, ld, StoreC{S});
}

// Second pass: out[k1 + N1 k2] = sum_n2 S[k1][n2] W_N2^(n2 k2).
template <int DIR>
static int fft_pass2(pcs_handle* h, const float2* S, float2* out) {
    const int N1 = 1 << h->logN1, N2 = 1 << h->logN2;
    TileGeom g2{};
    g2.nvec = N1;
    g2.in_vs = N2; g2.in_es = 1; g2.out_vs = 1; g2.out_es = N1;
    g2.v_fast_in = 0; g2.v_fast_out = 1;
    g2.twiddle = 0; g2.inv_ntw = 0.f; g2.ntw_mask = 0;
    if (N1 / 16 < h->sm_count / 2) return launch_tile_narrow<DIR>(h, h->logN2, g2, LoadC{S}, StoreC{out});
    return launch_tile<DIR>(h, h->logN2, g2, LoadC{S}, StoreC{out});
}

// N-point transform out = FFT_DIR(load(.)) through d_scratch (two passes).
template <int DIR, typename Load>
static int fft_large(pcs_handle* h, Load ld, float2* out) {
    if (int rc = fft_pass1<DIR>(h, ld, h->d_scratch)) return rc;
    return fft_pass2<DIR>(h, h->d_scratch, out);
}

// ---- overlap-save launchers ---------------------------------------------------------------------
template <int LOGB, int G, bool FS>
static int launch_search_os_t(pcs_handle* h, const OsSearchParams& p, bool locate) {
    using S = FftShape<LOGB>;
    constexpr int NW = (S::T + 31) / 32, NBUF = 3;
    const size_t smem = (size_t)G * NBUF * S::WORK * sizeof(float2) + (size_t)G * p.M * NW * 3 * sizeof(float);
    auto kern = search_os_kernel<LOGB, G, FS, false>;
    auto kloc = search_os_kernel<LOGB, G, FS, true>;
    static size_t configured[PCS_MAX_DEVICES] = {};
    if (configured[h->cfg.device] < smem) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(kloc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[h->cfg.device] = smem;
    }
    if (locate) {        // one group per (bin, mask): the peak's sample offset inside the winning block
        const long long items = (long long)p.D * p.M;
        kloc<<<(int)((items + G - 1) / G), G * S::T, smem, h->stream>>>(p);
        h->launches++;
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    if (FS) {      // block spectra of the unrotated chunk, once per chunk
        const size_t bs_smem = (size_t)G * 2 * S::WORK * sizeof(float2);
        auto bs = block_spectra_kernel<LOGB, G>;
        static size_t bs_configured[PCS_MAX_DEVICES] = {};
        if (bs_configured[h->cfg.device] < bs_smem) {
            CUDA_TRY(cudaFuncSetAttribute(bs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bs_smem));
            bs_configured[h->cfg.device] = bs_smem;
        }
        bs<<<(p.nblk + G - 1) / G, G * S::T, bs_smem, h->stream>>>(p.x, p.tw, h->osb[h->cur_lane].xbs, p.N, p.nblk, p.V, p.Lpos);
        h->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    const long long items = (long long)p.nblk * p.D;
    const int grid = (int)((items + G - 1) / G);
    h->search_ctas = grid;
    h->search_smem = (int)smem;
    kern<<<grid, G * S::T, smem, h->stream>>>(p);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Factorised bank: block spectra (as for the shifted-filter form), then search_fb_kernel with G = max(1, 2048 / B) groups
// per CTA (three 128-thread CTAs per SM up to B = 2048).
template <int LOGB, int GBS, int G, int J, int CB>
static int launch_search_fb_tj(pcs_handle* h, const OsSearchParams& p) {
    using S = FftShape<LOGB>;
    constexpr int NW = (S::T + 31) / 32;
    const size_t bs_smem = (size_t)GBS * 2 * S::WORK * sizeof(float2);
    auto bs = block_spectra_kernel<LOGB, GBS>;
    static size_t bs_configured[PCS_MAX_DEVICES] = {};
    if (bs_configured[h->cfg.device] < bs_smem) {
        CUDA_TRY(cudaFuncSetAttribute(bs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bs_smem));
        bs_configured[h->cfg.device] = bs_smem;
    }
    bs<<<(p.nblk + GBS - 1) / GBS, GBS * S::T, bs_smem, h->stream>>>(p.x, p.tw, h->osb[h->cur_lane].xbs, p.N, p.nblk, p.V, p.Lpos);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    FbSearchParams q{};
    q.xbs = p.xbs; q.tw = p.tw; q.psum = p.psum; q.pmax = p.pmax;
    q.gbasis = h->d_fb_basis + (((size_t)h->bin_lo * h->fb_R) << LOGB);
    q.coef = h->d_fb_coef + (size_t)h->bin_lo * h->M * J;
    q.sel = h->d_fb_sel;
    q.N = p.N; q.D = p.D; q.M = p.M; q.nblk = p.nblk; q.V = p.V; q.Lpos = p.Lpos; q.R = h->fb_R; q.S = h->fb_S;
    const size_t smem = (size_t)G * ((size_t)3 * S::WORK + (CB ? 0 : (size_t)(q.R - 1) * S::B)) * sizeof(float2) +
                        (size_t)G * p.M * NW * 2 * sizeof(float) + (size_t)G * p.M * J * sizeof(float2);
    auto kern = search_fb_kernel<LOGB, G, J, CB>;
    static size_t configured[PCS_MAX_DEVICES] = {};
    if (configured[h->cfg.device] < smem) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[h->cfg.device] = smem;
    }
    const long long items = (long long)p.nblk * p.D;
    const int grid = (int)((items + G - 1) / G);
    h->search_ctas = grid;
    h->search_smem = (int)smem;
    kern<<<grid, G * S::T, smem, h->stream>>>(q);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}
template <int LOGB, int GBS, int G>
static int launch_search_fb_t(pcs_handle* h, const OsSearchParams& p) {
    switch (h->fb_J) {
        case 2: return h->fb_complete == 2   ? launch_search_fb_tj<LOGB, GBS, G, 2, 2>(h, p)
                       : h->fb_complete == 1 ? launch_search_fb_tj<LOGB, GBS, G, 2, 1>(h, p)
                                             : launch_search_fb_tj<LOGB, GBS, G, 2, 0>(h, p);
        case 3: return h->fb_complete == 2   ? launch_search_fb_tj<LOGB, GBS, G, 3, 2>(h, p)
                       : h->fb_complete == 1 ? launch_search_fb_tj<LOGB, GBS, G, 3, 1>(h, p)
                                             : launch_search_fb_tj<LOGB, GBS, G, 3, 0>(h, p);
        case 4: return launch_search_fb_tj<LOGB, GBS, G, 4, 0>(h, p);
    }
    return fail(PCS_ERR_INVALID, "factorised bank with %d segments", h->fb_J);
}
static size_t fb_smem_bytes(int logB, int R, int M) {
    const size_t B = (size_t)1 << logB, work = B + (B >> 4), T = B / 16, G = std::max<size_t>(1, 2048 / B), NW = (T + 31) / 32;
    return G * (3 * work + (size_t)(R - 1) * B) * sizeof(float2) + G * M * NW * 2 * sizeof(float) + G * M * PCS_FB_MAX_SEG * sizeof(float2);
}
static int launch_search_fb(pcs_handle* h, const OsSearchParams& p) {
    switch (h->logB) {
        case 10: return launch_search_fb_t<10, 4, 2>(h, p);
        case 11: return launch_search_fb_t<11, 2, 1>(h, p);
        case 12: return launch_search_fb_t<12, 1, 1>(h, p);
    }
    return fail(PCS_ERR_INVALID, "factorised bank with 2^%d-point blocks", h->logB);
}

template <int LOGB, int G>
static int launch_demod_os_t(pcs_handle* h, const OsDemodParams& p) {
    using S = FftShape<LOGB>;
    const size_t smem = (size_t)G * 3 * S::WORK * sizeof(float2);
    auto kern = demod_os_kernel<LOGB, G>;
    static size_t configured[PCS_MAX_DEVICES] = {};
    if (configured[h->cfg.device] < smem) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[h->cfg.device] = smem;
    }
    const int grid = (p.nblk + G - 1) / G;
    kern<<<grid, G * S::T, smem, h->stream>>>(p);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int launch_search_os(pcs_handle* h, const OsSearchParams& p, bool locate) {
    switch (h->logB) {
        case 9: return p.xbs ? launch_search_os_t<9, 8, true>(h, p, locate) : launch_search_os_t<9, 8, false>(h, p, locate);
        case 10: return p.xbs ? launch_search_os_t<10, 4, true>(h, p, locate) : launch_search_os_t<10, 4, false>(h, p, locate);
        case 11: return p.xbs ? launch_search_os_t<11, 2, true>(h, p, locate) : launch_search_os_t<11, 2, false>(h, p, locate);
        case 12: return p.xbs ? launch_search_os_t<12, 1, true>(h, p, locate) : launch_search_os_t<12, 1, false>(h, p, locate);
        case 13: return p.xbs ? launch_search_os_t<13, 1, true>(h, p, locate) : launch_search_os_t<13, 1, false>(h, p, locate);
    }
    return fail(PCS_ERR_INVALID, "unsupported overlap-save block 2^%d", h->logB);
}
static int launch_demod_os(pcs_handle* h, const OsDemodParams& p) {
    switch (h->logB) {
        case 9: return launch_demod_os_t<9, 8>(h, p);
        case 10: return launch_demod_os_t<10, 4>(h, p);
        case 11: return launch_demod_os_t<11, 2>(h, p);
        case 12: return launch_demod_os_t<12, 1>(h, p);
        case 13: return launch_demod_os_t<13, 1>(h, p);
    }
    return fail(PCS_ERR_INVALID, "unsupported overlap-save block 2^%d", h->logB);
}
static int groups_for(int logB) { return logB == 9 ? 8 : logB == 10 ? 4 : logB == 11 ? 2 : 1; }

// ---- planning ---------------------------------------------------------------------------------------
// Time support of the filters: g_m = IFFT(Mk_m) has taps at n = 0..Lpos and n = -Lneg..-1 (circular).
static int measure_support(pcs_handle* h, int* Lpos, int* Lneg) {
    const int N = h->N, M = h->M;
    std::vector<float2> g((size_t)N);
    int lp = 0, ln = 0;
    for (int m = 0; m < M; ++m) {
        if (int rc = fft_large<+1>(h, LoadC{h->d_masks + (size_t)m * N}, h->d_X)) return rc;
        CUDA_TRY(cudaMemcpyAsync(g.data(), h->d_X, sizeof(float2) * N, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        double maxp = 0;
        for (int n = 0; n < N; ++n) maxp = std::max(maxp, (double)g[n].x * g[n].x + (double)g[n].y * g[n].y);
        const double thr = 1e-10 * maxp;   // 1e-5 in amplitude: far above complex64 rounding, far below any tap
        for (int n = N / 2 - 1; n > lp; --n)
            if ((double)g[n].x * g[n].x + (double)g[n].y * g[n].y > thr) { lp = n; break; }
        for (int k = N / 2; k > ln; --k)
            if ((double)g[N - k].x * g[N - k].x + (double)g[N - k].y * g[N - k].y > thr) { ln = k; break; }
    }
    *Lpos = lp;
    *Lneg = ln;
    return 0;
}

// Working set of one in-flight generic search (see pcs_handle::osb).
static int alloc_os_lane(pcs_handle* h, int lane) {
    pcs_handle::OsBufs& b = h->osb[lane];
    if (b.psum) return 0;
    const size_t np = (size_t)h->D * h->M * h->nblk;
    if (int rc = dev_alloc(h, &b.psum, np)) return rc;
    if (int rc = dev_alloc(h, &b.pmax, np)) return rc;
    if (int rc = dev_alloc(h, &b.wblk, (size_t)h->D * h->M)) return rc;
    if (h->os_fs)
        if (int rc = dev_alloc(h, &b.xbs, (size_t)h->nblk << h->logB)) return rc;
    return 0;
}

static int plan_overlap_save(pcs_handle* h, const float* masks_host) {
    const int N = h->N, M = h->M;
    const int L = h->Lpos + h->Lneg + 1;
    int best = 0;
    double best_cost = 1e30;
    const int lo = 9, hi = std::min(13, h->logN);
    for (int lb = lo; lb <= hi; ++lb) {
        const int B = 1 << lb, V = B - L + 1;
        if (V < B / 2) continue;
        const double cost = (double)lb * B / V;
        if (cost < best_cost * 0.97) { best_cost = cost; best = lb; }   // prefer the smaller block on near ties
    }
    if (h->cfg.log2_block && h->cfg.log2_block != 8) {
        best = h->cfg.log2_block;
        if (best < lo || best > hi || (1 << best) - L + 1 < 1)
            return fail(PCS_ERR_INVALID, "log2_block=%d cannot hold a filter support of %d taps", best, L);
    }
    if (!best) return PCS_ERR_INVALID;   // caller decides (AUTO falls back to FULL)
    h->logB = best;
    h->G = groups_for(best);
    const int B = 1 << best;
    h->V = B - L + 1;
    h->nblk = (N + h->V - 1) / h->V;
    // B-point filter spectra: Mk[m][k * N/B] * (N/B)
    const int dec = N / B;
    const float scale = (float)dec;
    std::vector<float2> gb((size_t)M * B);
    const float2* mk = reinterpret_cast<const float2*>(masks_host);
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < B; ++k) {
            const float2 v = mk[(size_t)m * N + (size_t)k * dec];
            gb[(size_t)m * B + k] = make_float2(v.x * scale, v.y * scale);
        }
    if (int rc = dev_alloc(h, &h->d_gb, gb.size())) return rc;
    CUDA_TRY(cudaMemcpyAsync(h->d_gb, gb.data(), sizeof(float2) * gb.size(), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    // Shifted-filter form (see plan_fast256) for the generic kernel, when the 256-point plan will not take the search
    // and the per-bin spectra stay modest (D * M * B * 8 bytes: 8 MiB for CC11xx); reserved[2] != 0 keeps the rotate form.
    const int L256 = 256 - L + 1;
    const bool fast256_will_run = (h->cfg.log2_block == 0 || h->cfg.log2_block == 8) && L256 >= 128 && N >= 4096;
    const size_t gs_elems = (size_t)h->D * M * B;
    const bool fs_form = h->cfg.reserved[2] == 0 || h->cfg.reserved[2] == 3;     // 3 = shifted filters, bank never factorised
    h->os_fs = fs_form && !fast256_will_run && gs_elems * sizeof(float2) <= ((size_t)256 << 20);
    if (int rc = alloc_os_lane(h, 0)) return rc;
    if (h->os_fs) {
        if (int rc = dev_alloc(h, &h->d_gs_os, gs_elems)) return rc;
        shifted_filters_kernel<<<(unsigned)((gs_elems + 255) / 256), 256, 0, h->stream>>>(h->d_masks, h->d_shifts, h->d_gs_os, N,
                                                                                          best, h->D, M);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    // Factorised bank (bank_factor.cu): R transforms per (bin, block) item instead of M when the M filters are combinations
    // of R < M basis segments (the reference's FSK-2 / CC11xx bank: R = 2, M = 8).
    if (h->os_fs && h->cfg.reserved[2] == 0 && best >= 10 && best <= 12 && M >= 2) {
        std::vector<int32_t> sel((size_t)M * PCS_FB_MAX_SEG);
        std::vector<float> coef((size_t)2 * h->D * M * PCS_FB_MAX_SEG), spec(((size_t)2 * h->D * PCS_FB_MAX_BASIS) << best);
        int32_t fS = 0, fJ = 0, fR = 0;
        if (int rc = pcs_factorise_bank(masks_host, N, M, h->Lpos, h->Lneg, h->h_shifts.data(), h->D, best, &fS, &fJ, &fR, sel.data(),
                                        coef.data(), spec.data()))
            return rc;
        int max_smem = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->cfg.device));
        if (fR > 0 && fb_smem_bytes(best, fR, M) <= (size_t)max_smem) {
            const size_t nb = ((size_t)h->D * fR) << best, nc = (size_t)h->D * M * fJ, ns = (size_t)M * fJ;
            // which form of search_fb_kernel the bank can take; a complete binary bank gets its tables in code order
            int32_t form = 1;
            if (int rc = pcs_bank_code_order(M, fJ, fR, h->D, sel.data(), coef.data(), (h->cfg.reserved[0] & 4) ? 0 : 1, &form)) return rc;
            h->fb_complete = form - 1;
            if (int rc = dev_alloc(h, &h->d_fb_basis, nb)) return rc;
            if (int rc = dev_alloc(h, &h->d_fb_coef, nc)) return rc;
            if (int rc = dev_alloc(h, &h->d_fb_sel, ns)) return rc;
            CUDA_TRY(cudaMemcpyAsync(h->d_fb_basis, spec.data(), sizeof(float2) * nb, cudaMemcpyHostToDevice, h->stream));
            CUDA_TRY(cudaMemcpyAsync(h->d_fb_coef, coef.data(), sizeof(float2) * nc, cudaMemcpyHostToDevice, h->stream));
            CUDA_TRY(cudaMemcpyAsync(h->d_fb_sel, sel.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, h->stream));
            CUDA_TRY(cudaStreamSynchronize(h->stream));
            h->fb = true; h->fb_S = fS; h->fb_J = fJ; h->fb_R = fR;
        }
    }
    return 0;
}

// Working set of one in-flight shifted-filter search (see pcs_handle::fsb).
static int alloc_fs256_lane(pcs_handle* h, int lane) {
    pcs_handle::Fs256Bufs& b = h->fsb[lane];
    if (b.xbs) return 0;
    const size_t np = (size_t)h->nblk256 * h->D * h->M;
    if (int rc = dev_alloc(h, &b.xbs, (size_t)h->nblk256 * 128)) return rc;
    if (int rc = dev_alloc(h, &b.psum, np)) return rc;
    if (int rc = dev_alloc(h, &b.pmax, np)) return rc;
    if (int rc = dev_alloc(h, &b.bin_count, (size_t)h->D + 1)) return rc;
    b.bins_done = b.bin_count + h->D;
    CUDA_TRY(cudaMemsetAsync(b.bin_count, 0, sizeof(unsigned int) * ((size_t)h->D + 1), h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

// Search plan for filters short enough for 256-point blocks (at least half of every block is valid output).
static int plan_fast256(pcs_handle* h, const float* masks_host) {
    const int N = h->N, M = h->M, D = h->D;
    const int L = h->Lpos + h->Lneg + 1;
    const int V = 256 - L + 1;
    if (V < 128 || N < 4096) return PCS_ERR_INVALID;
    h->V256 = V;
    h->nblk256 = (N + V - 1) / V;
    const int dec = N / 256;
    const float scale = (float)dec;
    const float2* mk = reinterpret_cast<const float2*>(masks_host);
    std::vector<float4> gp((size_t)M * 128);
    for (int m = 0; m < M; ++m)
        for (int rr = 0; rr < 8; ++rr)
            for (int t = 0; t < 16; ++t) {
                const float2 a = mk[(size_t)m * N + (size_t)(t + 16 * (2 * rr)) * dec];
                const float2 b = mk[(size_t)m * N + (size_t)(t + 16 * (2 * rr + 1)) * dec];
                gp[(size_t)m * 128 + rr * 16 + t] = make_float4(a.x * scale, a.y * scale, b.x * scale, b.y * scale);
            }
    if (int rc = dev_alloc(h, &h->d_gperm, gp.size())) return rc;
    CUDA_TRY(cudaMemcpyAsync(h->d_gperm, gp.data(), sizeof(float4) * gp.size(), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    const size_t np = (size_t)h->nblk256 * D * M;
    h->fast256 = true;
    // Shifted-filter form (default): per-bin filter spectra gathered once from the protocol's spectra, block spectra of
    // the unrotated chunk once per chunk.  reserved[2]: 0 = this form, 1 / 2 = rotate-the-chunk kernel (block spectrum
    // in shared memory / registers), kept as comparison variants.  More than 16 masks keep the rotate kernel (the
    // bin's spectra would not leave room for four CTAs per SM).
    if (h->cfg.reserved[2] == 0 && M <= 16) {
        if (int rc = alloc_fs256_lane(h, 0)) return rc;
        if (int rc = dev_alloc(h, &h->d_gs, (size_t)D * M * 128)) return rc;
        const long long n = (long long)D * M * 128;
        shifted_filters256_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->d_masks, h->d_shifts, h->d_gs, N, D, M);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        h->fs_items = h->cfg.reserved[1] >> 8;      // 0 = choose per launch
        h->fs256 = true;
        return 0;
    }
    if (int rc = dev_alloc(h, &h->d_psum256, np)) return rc;
    if (int rc = dev_alloc(h, &h->d_pmax256, np)) return rc;
    if (int rc = dev_alloc(h, &h->d_part_sum, (size_t)PCS_RED_SLICES * D * M)) return rc;
    if (int rc = dev_alloc(h, &h->d_part_max, (size_t)PCS_RED_SLICES * D * M)) return rc;
    if (int rc = dev_alloc(h, &h->d_part_blk, (size_t)PCS_RED_SLICES * D * M)) return rc;
    if (int rc = dev_alloc(h, &h->d_done, (size_t)1)) return rc;
    CUDA_TRY(cudaMemsetAsync(h->d_done, 0, sizeof(unsigned int), h->stream));
    return 0;
}

template <int MW>
static void launch_parseval_energy(pcs_handle* h, const float* W, const int* shifts, int Dl, int d_per_warp, int nslice) {
    const int ntile = h->N >> 8;
    const long long warps = (long long)ntile * nslice;
    parseval_energy_kernel<MW><<<(unsigned)((warps + 7) / 8), 256, 0, h->stream>>>(h->d_PX, W, shifts, h->d_pvpart, h->N, Dl,
                                                                                    d_per_warp);
    h->launches++;
}

static int plan_parseval(pcs_handle* h, const float* masks_host) {
    const int N = h->N, M = h->M, D = h->D;
    h->pv_mw = h->cfg.sum_all_masks ? 1 : M;
    const float2* mk = reinterpret_cast<const float2*>(masks_host);
    std::vector<float> W((size_t)h->pv_mw * N);
    for (int k = 0; k < N; ++k) {
        double sum = 0;
        for (int m = 0; m < M; ++m) {
            const float2 v = mk[(size_t)m * N + k];
            const double p = (double)v.x * v.x + (double)v.y * v.y;
            if (h->pv_mw == 1) sum += p; else W[(size_t)m * N + k] = (float)p;
        }
        if (h->pv_mw == 1) W[k] = (float)sum;
    }
    if (int rc = dev_alloc(h, &h->d_W, W.size())) return rc;
    CUDA_TRY(cudaMemcpyAsync(h->d_W, W.data(), sizeof(float) * W.size(), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (int rc = dev_alloc(h, &h->d_PX, (size_t)N)) return rc;
    if (int rc = dev_alloc(h, &h->d_pvpart, (size_t)(N >> 8) * D * std::min(h->pv_mw, 8))) return rc;
    h->parseval = true;
    return 0;
}

// Buffers the non-sharded tail of a chunk works in (chunk-spectrum pass 1, demod magnitudes, timing spectrum, result block).
// The handle owns one set; the streaming engine adds a second one so that the tails of two owned chunks can overlap.
struct TailSet {
    float2 *scratch = nullptr, *scratch2 = nullptr, *Pf = nullptr;
    float *ymag = nullptr, *p = nullptr;
    unsigned char* rblock = nullptr;
};
static void tail_get(const pcs_handle* h, TailSet* t) {
    t->scratch = h->d_scratch; t->scratch2 = h->d_scratch2; t->Pf = h->d_Pf; t->ymag = h->d_ymag; t->p = h->d_p;
    t->rblock = h->d_rblock;
}
static void tail_set(pcs_handle* h, const TailSet& t) {
    h->d_scratch = t.scratch; h->d_scratch2 = t.scratch2; h->d_Pf = t.Pf; h->d_ymag = t.ymag; h->d_p = t.p;
    h->d_rblock = t.rblock;
    h->d_res = reinterpret_cast<DevResult*>(t.rblock + h->rb_off[0]);
    h->d_E = reinterpret_cast<float*>(t.rblock + h->rb_off[1]);
    h->d_sym = reinterpret_cast<int*>(t.rblock + h->rb_off[2]);
    h->d_centre = reinterpret_cast<int*>(t.rblock + h->rb_off[3]);
    h->d_mag = reinterpret_cast<float*>(t.rblock + h->rb_off[4]);
    h->d_sigwin = reinterpret_cast<float2*>(t.rblock + h->rb_off[5]);
    h->d_noisewin = reinterpret_cast<float2*>(t.rblock + h->rb_off[6]);
}
static int tail_alloc(pcs_handle* h, TailSet* t) {
    const size_t N = (size_t)h->N;
    if (int rc = dev_alloc(h, &t->scratch, N)) return rc;
    if (int rc = dev_alloc(h, &t->scratch2, N)) return rc;
    if (int rc = dev_alloc(h, &t->Pf, N)) return rc;
    if (int rc = dev_alloc(h, &t->ymag, (size_t)h->M * N)) return rc;
    if (int rc = dev_alloc(h, &t->p, N)) return rc;
    if (int rc = dev_alloc(h, &t->rblock, h->rb_bytes)) return rc;
    CUDA_TRY(cudaMemsetAsync(t->rblock, 0, h->rb_bytes, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

// ---- public API -----------------------------------------------------------------------------------
extern "C" {

const char* pcs_last_error(void) { return g_last_error.c_str(); }
int pcs_abi_version(void) { return PCS_ABI_VERSION; }

static void shard_destroy(pcs_handle* h);

int pcs_destroy(pcs_handle* h) {
    if (!h) return PCS_OK;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->side) cudaStreamSynchronize(h->side);
    for (void* p : h->dev_allocs) cudaFree(p);
    void* pinned[] = {h->h_x, h->h_sigwin, h->h_noisewin, h->h_res, h->h_E, h->h_mag, h->h_sym, h->h_centre, h->h_thr_bits};
    for (void* p : pinned)
        if (p) cudaFreeHost(p);
    shard_destroy(h);
    if (h->gexec) cudaGraphExecDestroy(h->gexec);
    if (h->side) cudaStreamDestroy(h->side);
    for (cudaEvent_t e : {h->ev_fork, h->ev_est, h->ev_side})
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev)
        if (e) cudaEventDestroy(e);
    if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return PCS_OK;
}

static int create_impl(pcs_handle* h, const pcs_config* cfg, const int32_t* shifts, const float* masks) {
    h->cfg = *cfg;
    if (h->cfg.snr_window <= 0) h->cfg.snr_window = 5;
    h->graph_enabled = !(cfg->reserved[0] & 1);
    const int N = cfg->nfft;
    if (N < (1 << 12) || N > (1 << 24) || (N & (N - 1)))
        return fail(PCS_ERR_INVALID, "nfft must be a power of two in [2^12, 2^24], got %d", N);
    if (cfg->num_masks < 1 || cfg->num_masks > 32)
        return fail(PCS_ERR_INVALID, "num_masks must be in [1, 32], got %d", cfg->num_masks);
    if (cfg->num_dopplers < 1 || cfg->element_offset < 0 || cfg->element_offset > 1)
        return fail(PCS_ERR_INVALID, "bad Doppler grid (%d bins, offset %d)", cfg->num_dopplers, cfg->element_offset);
    if (cfg->window_width < 1 || cfg->window_width > 63 || !(cfg->window_width & 1))
        return fail(PCS_ERR_INVALID, "window_width must be odd and in [1, 63], got %d", cfg->window_width);
    if (cfg->samples_per_sym < 2) return fail(PCS_ERR_INVALID, "samples_per_sym must be >= 2");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(PCS_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
    if (cfg->device >= PCS_MAX_DEVICES) return fail(PCS_ERR_INVALID, "device index %d too large", cfg->device);
    if (cfg->device < 0 || cfg->device >= ndev) return fail(PCS_ERR_INVALID, "device %d out of range (%d devices)", cfg->device, ndev);
    CUDA_TRY(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) return fail(PCS_ERR_NO_DEVICE, "device %s is sm_%d%d; this build targets sm_100a only", prop.name, prop.major, prop.minor);
    h->sm_count = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_est, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_side, cudaEventDisableTiming));
    h->N = N;
    h->logN = ilog2(N);
    h->logN1 = h->logN / 2;
    h->logN2 = h->logN - h->logN1;
    h->D = cfg->num_dopplers + cfg->element_offset;
    h->M = cfg->num_masks;
    h->bin_lo = 0;
    h->bin_hi = h->D;
    const int D = h->D, M = h->M;
    h->spsym_min = cfg->samples_per_sym / 2;                               // dem_base:104
    if (h->spsym_min < 1) h->spsym_min = 1;
    h->max_sym = N / h->spsym_min;                                         // dem_base:468
    h->i_low = (int)((double)N / (0.9 * (double)cfg->samples_per_sym));    // dem_base:508-512
    h->i_high = (int)((double)N / (1.1 * (double)cfg->samples_per_sym));
    if (h->i_low > N / 2) h->i_low = N / 2;
    for (int d = 0; d < D; ++d)
        if (shifts[d] < 0 || shifts[d] >= N) return fail(PCS_ERR_INVALID, "shift[%d]=%d outside [0, nfft)", d, shifts[d]);
    {   // largest window computeSNR can ask for: neighbouring bins' spacing + 2 * half width
        int cap = 2 * h->cfg.snr_window;
        for (int d = cfg->element_offset; d + 1 < D; ++d)
            cap = std::max(cap, ((shifts[d + 1] - shifts[d]) & (N - 1)) + 2 * h->cfg.snr_window);
        h->win_cap = std::min(cap, PCS_WINDOW_MAX);
    }

    if (int rc = dev_alloc(h, &h->d_x, (size_t)N)) return rc;
    if (int rc = dev_alloc(h, &h->d_X, (size_t)N)) return rc;
    if (int rc = dev_alloc(h, &h->d_scratch, (size_t)N)) return rc;
    if (int rc = dev_alloc(h, &h->d_scratch2, (size_t)N)) return rc;
    if (int rc = dev_alloc(h, &h->d_Pf, (size_t)N)) return rc;
    if (int rc = dev_alloc(h, &h->d_masks, (size_t)M * N)) return rc;
    if (int rc = dev_alloc(h, &h->d_shifts, (size_t)D)) return rc;
    if (int rc = dev_alloc(h, &h->d_Efull, (size_t)D * M)) return rc;
    if (int rc = dev_alloc(h, &h->d_peakv, (size_t)D * M)) return rc;
    if (int rc = dev_alloc(h, &h->d_peako, (size_t)D * M)) return rc;
    h->tab_E = h->d_Efull; h->tab_pv = h->d_peakv; h->tab_po = h->d_peako;
    if (int rc = dev_alloc(h, &h->d_ymag, (size_t)M * N)) return rc;
    if (int rc = dev_alloc(h, &h->d_p, (size_t)N)) return rc;
    {   // result block: res | E | sym | centre | mag | sigwin | noisewin
        const size_t sizes[7] = {sizeof(DevResult), sizeof(float) * D * M, sizeof(int) * h->max_sym, sizeof(int) * h->max_sym,
                                 sizeof(float) * h->max_sym, sizeof(float2) * h->win_cap, sizeof(float2) * h->win_cap};
        size_t off = 0;
        for (int i = 0; i < 7; ++i) {
            h->rb_off[i] = off;
            off += (sizes[i] + 255) / 256 * 256;
        }
        h->rb_bytes = off;
        if (int rc = dev_alloc(h, &h->d_rblock, off)) return rc;
        h->d_res = reinterpret_cast<DevResult*>(h->d_rblock + h->rb_off[0]);
        h->d_E = reinterpret_cast<float*>(h->d_rblock + h->rb_off[1]);
        h->d_sym = reinterpret_cast<int*>(h->d_rblock + h->rb_off[2]);
        h->d_centre = reinterpret_cast<int*>(h->d_rblock + h->rb_off[3]);
        h->d_mag = reinterpret_cast<float*>(h->d_rblock + h->rb_off[4]);
        h->d_sigwin = reinterpret_cast<float2*>(h->d_rblock + h->rb_off[5]);
        h->d_noisewin = reinterpret_cast<float2*>(h->d_rblock + h->rb_off[6]);
    }
    CUDA_TRY(cudaMemsetAsync(h->d_res, 0, sizeof(DevResult), h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d_sym, 0, sizeof(int) * h->max_sym, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d_centre, 0, sizeof(int) * h->max_sym, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d_mag, 0, sizeof(float) * h->max_sym, h->stream));

    CUDA_TRY(cudaHostAlloc((void**)&h->h_x, sizeof(float2) * N, cudaHostAllocDefault));
    memset(h->h_x, 0, sizeof(float2) * N);
    CUDA_TRY(cudaHostAlloc((void**)&h->h_res, sizeof(DevResult), cudaHostAllocDefault));
    memset(h->h_res, 0, sizeof(DevResult));
    CUDA_TRY(cudaHostAlloc((void**)&h->h_E, sizeof(float) * D * M, cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc((void**)&h->h_sym, sizeof(int) * h->max_sym, cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc((void**)&h->h_centre, sizeof(int) * h->max_sym, cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc((void**)&h->h_mag, sizeof(float) * h->max_sym, cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc((void**)&h->h_sigwin, sizeof(float2) * PCS_WINDOW_MAX, cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc((void**)&h->h_noisewin, sizeof(float2) * PCS_WINDOW_MAX, cudaHostAllocDefault));

    CUDA_TRY(cudaMemcpyAsync(h->d_masks, masks, sizeof(float2) * (size_t)M * N, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_shifts, shifts, sizeof(int) * D, cudaMemcpyHostToDevice, h->stream));
    h->h_shifts.assign(shifts, shifts + D);
    CUDA_TRY(cudaStreamSynchronize(h->stream));

    if (int rc = measure_support(h, &h->Lpos, &h->Lneg)) return rc;
    const int want = cfg->path;
    if (want == PCS_PATH_FULL)
        return fail(PCS_ERR_INVALID, "PCS_PATH_FULL (Nfft-point inverse transforms) is not built: overlap-save gives the same "
                                     "numbers for every filter support up to 2^13 taps");
    if (want == PCS_PATH_AUTO || want == PCS_PATH_OVERLAP_SAVE || want == PCS_PATH_PARSEVAL) {
        int rc = plan_overlap_save(h, masks);
        if (rc == 0) {
            h->path = PCS_PATH_OVERLAP_SAVE;
        } else if (want == PCS_PATH_OVERLAP_SAVE) {
            return fail(PCS_ERR_INVALID, "filter time support (%d + %d taps) too long for the overlap-save path",
                        h->Lpos, h->Lneg);
        } else if (rc != PCS_ERR_INVALID) {
            return rc;
        }
    }
    if (h->path == 0)
        return fail(PCS_ERR_INVALID, "path %d is not available in this build for a filter support of %d + %d taps",
                    want, h->Lpos, h->Lneg);
    if (want == PCS_PATH_PARSEVAL) {
        if (int rc = plan_parseval(h, masks)) return rc;
        h->path = PCS_PATH_PARSEVAL;
    }
    if (cfg->log2_block == 0 || cfg->log2_block == 8) {
        int rc = plan_fast256(h, masks);
        if (rc != 0 && rc != PCS_ERR_INVALID) return rc;
        if (rc != 0 && cfg->log2_block == 8)
            return fail(PCS_ERR_INVALID, "log2_block=8 cannot hold a filter support of %d taps", h->Lpos + h->Lneg + 1);
    }
    return PCS_OK;
}

int pcs_create(const pcs_config* cfg, const int32_t* shifts, const float* masks, pcs_handle** out) {
    if (!cfg || !shifts || !masks || !out) return fail(PCS_ERR_INVALID, "null argument");
    if (cfg->abi_version != PCS_ABI_VERSION)
        return fail(PCS_ERR_INVALID, "ABI version mismatch: caller %d, library %d", cfg->abi_version, PCS_ABI_VERSION);
    pcs_handle* h = new pcs_handle();
    int rc = create_impl(h, cfg, shifts, masks);
    if (rc != PCS_OK) {
        std::string keep = g_last_error;
        pcs_destroy(h);
        g_last_error = keep;
        *out = nullptr;
        return rc;
    }
    *out = h;
    return PCS_OK;
}

void* pcs_host_buffer(pcs_handle* h) { return h ? (void*)h->h_x : nullptr; }

// Page-lock / release a range of the caller's own memory (e.g. the sample ring a receiver thread writes into) so that chunks
// can be copied to HBM straight from it: the caller's fill of the handle's pinned buffer (demodulator_process.py:287) is
// 8 * nfft bytes of host memcpy per chunk, more than the H2D copy itself costs.
int pcs_host_register(void* ptr, uint64_t bytes) {
    if (!ptr || !bytes) return fail(PCS_ERR_INVALID, "null or empty range");
    CUDA_TRY(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
    return PCS_OK;
}

int pcs_host_unregister(void* ptr) {
    if (!ptr) return fail(PCS_ERR_INVALID, "null pointer");
    CUDA_TRY(cudaHostUnregister(ptr));
    return PCS_OK;
}

int pcs_set_host_source(pcs_handle* h, const void* chunk) {
    if (!h) return fail(PCS_ERR_INVALID, "null handle");
    if (!chunk) { h->h_src = nullptr; return PCS_OK; }
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    // both ends of the chunk must be page-locked: a pageable source would make the "asynchronous" copy a staged, blocking one
    const char* ends[2] = {reinterpret_cast<const char*>(chunk), reinterpret_cast<const char*>(chunk) + sizeof(float2) * (size_t)h->N - 1};
    for (const char* e : ends) {
        cudaPointerAttributes at{};
        const cudaError_t rc = cudaPointerGetAttributes(&at, e);
        if (rc != cudaSuccess || at.type != cudaMemoryTypeHost) {
            cudaGetLastError();
            return fail(PCS_ERR_INVALID, "pcs_set_host_source: the chunk is not in page-locked host memory (pcs_host_register it)");
        }
    }
    h->h_src = reinterpret_cast<const float2*>(chunk);
    return PCS_OK;
}
int32_t pcs_max_symbols(const pcs_handle* h) { return h ? h->max_sym : 0; }
int64_t pcs_launch_count(const pcs_handle* h) { return h ? h->launches : 0; }
uint64_t pcs_stream(const pcs_handle* h) { return h ? (uint64_t)(uintptr_t)h->stream : 0; }

int pcs_get_plan(const pcs_handle* h, pcs_plan_info* info) {
    if (!h || !info) return fail(PCS_ERR_INVALID, "null argument");
    info->path = h->path;
    info->log2_block = h->fast256 ? 8 : h->logB;
    info->valid_per_block = h->fast256 ? h->V256 : h->V;
    info->num_blocks = h->fast256 ? h->nblk256 : h->nblk;
    info->support_pos = h->Lpos;
    info->support_neg = h->Lneg;
    const int gk = h->cfg.reserved[1] & 0xff, g256 = gk == 16 ? 16 : gk == 4 ? 4 : 8;
    info->groups_per_cta = h->fast256 ? g256 : h->G;
    info->search_ctas = h->fs256 ? h->search_ctas          // of the last launch (the tiling adapts to the bin range)
                        : h->fast256 ? (int)(((long long)h->nblk256 * h->D + g256 - 1) / g256)
                                     : (int)(((long long)h->nblk * h->D + h->G - 1) / h->G);
    info->search_smem_bytes = h->search_smem;
    info->sm_count = h->sm_count;
    info->device_bytes = h->dev_bytes;
    return PCS_OK;
}

int pcs_get_bank_factor(const pcs_handle* h, int32_t out[4]) {
    if (!h || !out) return fail(PCS_ERR_INVALID, "null argument");
    out[0] = h->fb && !h->fast256 ? 1 + h->fb_complete : 0; out[1] = h->fb_S; out[2] = h->fb_J; out[3] = h->fb_R;
    return PCS_OK;
}

// ---- enqueue helpers (no synchronisation) -------------------------------------------------------------
// a4 (dem_base:557).  Only the spectrum bins computeSNR averages are ever consumed per chunk (the search works on the
// time-domain chunk), so the hot path runs pass 1 of the four-step transform here and spectrum_bins_kernel finishes
// just those bins after the estimate; pcs_get_spectrum completes the full transform on demand.
static int enqueue_spectrum(pcs_handle* h) {
    StageTimer t(h, PCS_STAGE_SPECTRUM);
    h->spectrum_full = false;
    return fft_pass1<-1>(h, LoadC{h->d_x_cur}, h->d_scratch);
}

static int enqueue_estimate(pcs_handle* h);

// Items (bin, block) per CTA of the shifted-filter search.  Large searches take 64 (measured on C2: 0.761 ms against 0.778 /
// 0.781 ms for 56 / 112, profiles/r02_kernel_variants.md: CTAs do not run in lockstep waves, so finer CTAs only add set-up).
// A rank's slice of the bins would leave the last wave mostly empty with that (C2 on 8 GPUs: 628 CTAs on 592 slots), so
// small searches pick the multiple of the group count G (a CTA does ceil(items / G) rounds) that minimises rounds x waves,
// with a small per-CTA set-up charge (twiddles + the bin's M x 2 KB filter spectra).
static int choose_fs_items(long long items, int G, int sm_count, int cap) {
    const long long slots = 4LL * sm_count;
    if (items >= 6 * 64 * slots && (cap <= 0 || cap >= 64)) return 64;
    int best = G;
    double best_cost = 1e300;
    const int top = cap > 0 ? std::max(G, std::min(128, cap)) : 128;
    for (int ipc = top / G * G; ipc >= G; ipc -= G) {          // descending: the larger CTA wins ties
        const long long ctas = (items + ipc - 1) / ipc, waves = (ctas + slots - 1) / slots;
        const double cost = (double)waves * ((double)ipc / G + 0.35);
        if (cost < best_cost) { best_cost = cost; best = ipc; }
    }
    return best;
}

// Search of the handle's bin range [bin_lo, bin_hi) into the tables tab_E / tab_pv / tab_po (rows bin_lo..bin_hi-1).
static int enqueue_search_local256(pcs_handle* h) {
    const int Dl = h->bin_hi - h->bin_lo, DM = Dl * h->M;
    const size_t row0 = (size_t)h->bin_lo * h->M;
    Os256Params p{};
    p.x = h->d_x_cur; p.gperm = h->d_gperm; p.shifts = h->d_shifts + h->bin_lo;
    p.psum = h->d_psum256; p.pmax = h->d_pmax256;
    p.N = h->N; p.D = Dl; p.M = h->M; p.nblk = h->nblk256; p.V = h->V256; p.Lpos = h->Lpos;
    p.invN = 1.0f / (float)h->N;
    const float2* twp = nullptr;
    if (int rc = get_twiddles(h, 8, &twp)) return rc;
    p.tw = twp;
    const int Gk = h->cfg.reserved[1] & 0xff;
    if (h->fs256) {
        // block spectra of the chunk, then ONE kernel: filter products, inverse transforms, |y|^2 sum / max per (bin,
        // block), and -- by the CTA that completes a bin -- the bin's fixed-order reduction, peak offset, table row and
        // (bin sharding) the arrival flag in the owner's exchange region
        pcs_handle::Fs256Bufs& lb = h->fsb[h->cur_lane];
        const float4* xbs = h->xbs_ext;
        h->xbs_ext = nullptr;
        if (!xbs) {
            StageTimer tb(h, PCS_STAGE_BLOCK_SPECTRA);
            block_spectra256_kernel<4><<<(p.nblk + 3) / 4, 64, 0, h->stream>>>(p.x, p.tw, lb.xbs, p.N, p.nblk, p.V, p.Lpos);
            h->launches++;
            CUDA_TRY(cudaGetLastError());
            xbs = lb.xbs;
        }
        StageTimer t(h, PCS_STAGE_SEARCH);
        Fs256Params q{};
        q.xbs = xbs; q.gs = h->d_gs + (size_t)h->bin_lo * h->M * 128; q.tw = p.tw; q.psum = lb.psum; q.pmax = lb.pmax;
        q.N = p.N; q.D = Dl; q.M = p.M; q.nblk = p.nblk; q.V = p.V; q.Lpos = p.Lpos;
        q.bin_count = lb.bin_count; q.bins_done = lb.bins_done;
        q.Efull = h->tab_E + row0; q.peak_val = h->tab_pv + row0; q.peak_off = h->tab_po + row0;
        q.arrival_flag = h->push_flag; q.arrival_value = h->push_value;
        q.ack_flag = h->push_flag ? h->push_ack : nullptr;
        if (h->push_flag) h->push_ack = nullptr;
        h->push_flag = nullptr;       // consumed: the search kernel raises the flag(s) itself
        const long long items = (long long)p.nblk * Dl;
        const int G = Gk == 16 ? 16 : Gk == 4 ? 4 : 8;
        q.items_per_cta = h->fs_items > 0 ? h->fs_items : choose_fs_items(items, G, h->sm_count, h->fs_items_cap);
        h->search_ctas = (int)((items + q.items_per_cta - 1) / q.items_per_cta);
        const size_t dyn = (size_t)p.M * 128 * sizeof(float4) + (size_t)G * 2 * p.M * 17 * sizeof(float);
        h->search_smem = (int)(G * 272 * sizeof(float2) + dyn);
        static size_t configured[PCS_MAX_DEVICES][4] = {};
        const bool occ20 = (h->cfg.reserved[0] & 2) && G == 8;   // tuning knob: 96-register build, 5 CTAs (20 warps) per SM
        const int gi = occ20 ? 3 : G == 16 ? 0 : G == 8 ? 1 : 2;
        auto kern = occ20 ? search_fs256_kernel<8, 20> : G == 16 ? search_fs256_kernel<16> : G == 8 ? search_fs256_kernel<8>
                                                                                            : search_fs256_kernel<4>;
        if (configured[h->cfg.device][gi] < dyn) {
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            configured[h->cfg.device][gi] = dyn;
        }
        kern<<<h->search_ctas, G * 16, dyn, h->stream>>>(q);
        h->launches++;
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    {
        StageTimer t(h, PCS_STAGE_SEARCH);
        const long long items = (long long)p.nblk * Dl;
        const int G = Gk == 16 ? 16 : Gk == 4 ? 4 : 8;   // groups per CTA (8: measured best)
        h->search_ctas = (int)((items + G - 1) / G);
        const size_t acc_bytes = (size_t)G * 2 * p.M * 17 * sizeof(float);
        h->search_smem = (int)(G * 272 * sizeof(float2) + acc_bytes);
        static size_t configured[PCS_MAX_DEVICES][6] = {};   // static + dynamic shared memory may exceed the 48 KB default
        const bool xbs = h->cfg.reserved[2] == 1 && G != 16;   // tuning knob: block spectrum in shared memory
        const int gi = (G == 16 ? 0 : G == 8 ? 1 : 2) + (xbs ? 3 : 0);
        auto kern = xbs ? (G == 8 ? search_os256_kernel<8, true> : search_os256_kernel<4, true>)
                        : (G == 16 ? search_os256_kernel<16, false> : G == 8 ? search_os256_kernel<8, false>
                                                                              : search_os256_kernel<4, false>);
        if (configured[h->cfg.device][gi] < acc_bytes) {
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)acc_bytes));
            configured[h->cfg.device][gi] = acc_bytes;
        }
        kern<<<h->search_ctas, G * 16, acc_bytes, h->stream>>>(p);
        h->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    StageTimer t2(h, PCS_STAGE_REDUCE);
    search_reduce256_kernel<<<dim3((DM + 31) / 32, PCS_RED_SLICES), dim3(32, 32), 0, h->stream>>>(
        p.psum, p.pmax, DM, p.nblk, h->d_part_sum, h->d_part_max, h->d_part_blk);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    peak_locate256_kernel<<<(DM + 15) / 16, 256, 0, h->stream>>>(p, h->d_part_sum, h->d_part_max, h->d_part_blk,
                                                                  h->tab_E + row0, h->tab_pv + row0, h->tab_po + row0,
                                                                  h->d_done, h->push_flag, h->push_value);
    h->push_flag = nullptr;       // consumed: the locate kernel raises the flag itself
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Block spectra of chunk `x` into `out` on `stream` (the streaming engine's ingest rank computes them once for every rank).
static int enqueue_block_spectra256(pcs_handle* h, cudaStream_t stream, const float2* x, float4* out) {
    const float2* twp = nullptr;
    if (int rc = get_twiddles(h, 8, &twp)) return rc;
    block_spectra256_kernel<4><<<(h->nblk256 + 3) / 4, 64, 0, stream>>>(x, twp, out, h->N, h->nblk256, h->V256, h->Lpos);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int enqueue_search_local_parseval(pcs_handle* h) {
    const int Dl = h->bin_hi - h->bin_lo, M = h->M, N = h->N;
    const size_t row0 = (size_t)h->bin_lo * M;
    StageTimer t(h, PCS_STAGE_SEARCH);
    if (!h->spectrum_full) {
        if (int rc = fft_large<-1>(h, LoadC{h->d_x_cur}, h->d_X)) return rc;
        h->spectrum_full = true;
    }
    abs2_kernel<<<(N + 255) / 256, 256, 0, h->stream>>>(h->d_X, h->d_PX, N);
    h->launches++;
    const int d_per_warp = 32, nslice = (Dl + d_per_warp - 1) / d_per_warp;
    const float scale = (float)N / 262144.0f;                     // kern:442 with Parseval's factor N
    for (int m0 = 0; m0 < h->pv_mw;) {
        const int rem = h->pv_mw - m0, mw = rem >= 8 ? 8 : rem >= 4 ? 4 : rem >= 2 ? 2 : 1;
        const float* W = h->d_W + (size_t)m0 * N;
        const int* sh = h->d_shifts + h->bin_lo;
        switch (mw) {
            case 8: launch_parseval_energy<8>(h, W, sh, Dl, d_per_warp, nslice); break;
            case 4: launch_parseval_energy<4>(h, W, sh, Dl, d_per_warp, nslice); break;
            case 2: launch_parseval_energy<2>(h, W, sh, Dl, d_per_warp, nslice); break;
            default: launch_parseval_energy<1>(h, W, sh, Dl, d_per_warp, nslice); break;
        }
        CUDA_TRY(cudaGetLastError());
        // the last batch also zero-fills the columns SUM mode leaves empty
        const bool last = m0 + mw >= h->pv_mw;
        parseval_reduce_kernel<<<(Dl * mw + 31) / 32, dim3(32, 32), 0, h->stream>>>(
            h->d_pvpart, N >> 8, Dl, mw, M, last ? M - (m0 + mw) : 0, scale, h->tab_E + row0 + m0, h->tab_pv + row0 + m0,
            h->tab_po + row0 + m0);
        h->launches++;
        CUDA_TRY(cudaGetLastError());
        m0 += mw;
    }
    return 0;
}

static int enqueue_search_local(pcs_handle* h) {
    if (h->parseval) return enqueue_search_local_parseval(h);
    if (h->fast256) return enqueue_search_local256(h);
    const int Dl = h->bin_hi - h->bin_lo;
    const size_t row0 = (size_t)h->bin_lo * h->M;
    OsSearchParams p{};
    p.x = h->d_x_cur; p.gb = h->d_gb; p.shifts = h->d_shifts + h->bin_lo;
    pcs_handle::OsBufs& lb = h->osb[h->cur_lane];
    p.xbs = lb.xbs;
    p.gs = h->d_gs_os ? h->d_gs_os + ((size_t)row0 << h->logB) : nullptr;
    p.psum = lb.psum + row0 * h->nblk; p.pmax = lb.pmax + row0 * h->nblk;
    p.wblk = lb.wblk + row0; p.peak_off = h->tab_po + row0;
    p.N = h->N; p.D = Dl; p.M = h->M; p.nblk = h->nblk; p.V = h->V; p.Lpos = h->Lpos;
    p.invN = 1.0f / (float)h->N;
    const float2* twp = nullptr;
    if (int rc = get_twiddles(h, h->logB, &twp)) return rc;
    p.tw = twp;
    if (int rc = get_pass_twiddles(h, h->logB, &twp)) return rc;
    p.twp = twp;
    {
        StageTimer t(h, PCS_STAGE_SEARCH);
        if (int rc = h->fb ? launch_search_fb(h, p) : launch_search_os(h, p, false)) return rc;
    }
    StageTimer t2(h, PCS_STAGE_REDUCE);
    const int DM = Dl * h->M;
    search_reduce_kernel<<<(DM * 32 + 255) / 256, 256, 0, h->stream>>>(p.psum, p.pmax, DM, h->nblk, h->tab_E + row0,
                                                                        h->tab_pv + row0, lb.wblk + row0);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return launch_search_os(h, p, true);
}

static int enqueue_search(pcs_handle* h) {
    if (int rc = enqueue_search_local(h)) return rc;
    return enqueue_estimate(h);
}

// findDopplerEst + shift interpolation + peak + SNR windows over the full [D][M] energy table (which, with bin
// sharding, the caller has all-gathered into d_Efull / d_peakv / d_peako beforehand).
static int enqueue_estimate_kernel(pcs_handle* h) {
    EstimateParams e{};
    e.Efull = h->tab_E; e.E = h->d_E; e.peak_val = h->tab_pv; e.peak_off = h->tab_po; e.shifts = h->d_shifts;
    e.res = h->d_res;
    e.D = h->D; e.M = h->M; e.N = h->N; e.num_dopplers = h->cfg.num_dopplers; e.element_offset = h->cfg.element_offset;
    e.sum_all = h->cfg.sum_all_masks ? 1 : 0; e.window_width = h->cfg.snr_window; e.window_cap = h->win_cap;
    estimate_kernel<<<1, 256, 0, h->stream>>>(e);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int enqueue_snr_bins(pcs_handle* h) {
    if (h->spectrum_full) {        // a full spectrum of this chunk exists (Parseval variant): just gather the windows
        spectrum_gather_kernel<<<(h->win_cap + 255) / 256, 256, 0, h->stream>>>(h->d_X, h->N, h->d_res, h->d_sigwin,
                                                                             h->d_noisewin);
        h->launches++;
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    const float2* tw2 = nullptr;
    if (int rc = get_twiddles(h, h->logN2, &tw2)) return rc;
    spectrum_bins_kernel<<<(2 * h->win_cap + 7) / 8, 256, 0, h->stream>>>(h->d_scratch, tw2, 1 << h->logN1, 1 << h->logN2,
                                                                          h->d_res, h->d_sigwin, h->d_noisewin);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int enqueue_estimate(pcs_handle* h) {
    if (h->spectrum_pending && !h->spectrum_full) {
        if (int rc = enqueue_spectrum(h)) return rc;
    }
    h->spectrum_pending = false;
    StageTimer t2(h, PCS_STAGE_ESTIMATE);
    if (int rc = enqueue_estimate_kernel(h)) return rc;
    return enqueue_snr_bins(h);
}

static int enqueue_timing_spectrum(pcs_handle* h) {
    // Pf = RFFT_N(p) (dem_base:721) restricted to the symbol-rate band findCodeRateAndPhase scans (dem_base:508-512)
    const int N1 = 1 << h->logN1, N2 = 1 << h->logN2;
    const int k2lo = h->i_high / N1, k2hi = std::max(h->i_low - 1, h->i_high) / N1, nk2 = k2hi - k2lo + 1;
    if (int rc = fft_pass1<-1>(h, LoadR{h->d_p}, h->d_scratch2)) return rc;
    if (nk2 > 16) return fft_pass2<-1>(h, h->d_scratch2, h->d_Pf);
    const float2* tw2 = nullptr;
    if (int rc = get_twiddles(h, h->logN2, &tw2)) return rc;
    band_dft_kernel<<<(N1 + 7) / 8, 256, 0, h->stream>>>(h->d_scratch2, tw2, N1, N2, k2lo, nk2, h->d_Pf);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int enqueue_demod_surface256(pcs_handle* h, int shift, bool want_complex) {
    Demod256Params q{};
    q.os.x = h->d_x_cur; q.os.gperm = h->d_gperm; q.os.N = h->N; q.os.M = h->M; q.os.nblk = h->nblk256; q.os.V = h->V256;
    q.os.Lpos = h->Lpos; q.os.D = 1; q.os.invN = 1.0f / (float)h->N;
    const float2* twp = nullptr;
    if (int rc = get_twiddles(h, 8, &twp)) return rc;
    q.os.tw = twp;
    q.res = h->d_res; q.ymag = h->d_ymag; q.p = h->d_p; q.ycplx = want_complex ? h->d_ycplx : nullptr;
    q.shift_override = shift;
    q.mask_lo = h->cfg.code_search_mask_offset; q.mask_hi = h->M - h->cfg.code_search_mask_offset;
    demod_os256_kernel<<<(h->nblk256 + 3) / 4, 64, 0, h->stream>>>(q);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int enqueue_demod(pcs_handle* h, int shift, bool want_complex) {
    OsDemodParams p{};
    p.x = h->d_x_cur; p.gb = h->d_gb; p.res = h->d_res; p.ymag = h->d_ymag; p.p = h->d_p;
    p.ycplx = want_complex ? h->d_ycplx : nullptr;
    p.N = h->N; p.M = h->M; p.nblk = h->nblk; p.V = h->V; p.Lpos = h->Lpos; p.shift_override = shift;
    p.mask_lo = h->cfg.code_search_mask_offset; p.mask_hi = h->M - h->cfg.code_search_mask_offset;
    p.invN = 1.0f / (float)h->N;
    const float2* twp = nullptr;
    if (int rc = get_twiddles(h, h->logB, &twp)) return rc;
    p.tw = twp;
    {
        StageTimer t(h, PCS_STAGE_DEMOD_SURFACE);
        if (h->fast256) {
            if (int rc = enqueue_demod_surface256(h, shift, want_complex)) return rc;
        } else {
            if (int rc = launch_demod_os(h, p)) return rc;
        }
    }
    if (want_complex) return 0;
    StageTimer t2(h, PCS_STAGE_TIMING_SYMBOLS);
    if (int rc = enqueue_timing_spectrum(h)) return rc;
    timing_kernel<<<1, 1024, 0, h->stream>>>(h->d_Pf, h->i_high, h->i_low - h->i_high, h->N, h->spsym_min, h->d_res);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    centres_kernel<<<(h->max_sym + 255) / 256, 256, 0, h->stream>>>(h->d_ymag, h->d_res, h->N, h->M, h->cfg.window_width,
                                                                  h->spsym_min, h->max_sym, h->d_sym, h->d_centre, h->d_mag);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int enqueue_fetch_windows(pcs_handle* h);
static int enqueue_fetch_search(pcs_handle* h) {
    CUDA_TRY(cudaMemcpyAsync(h->h_res, h->d_res, sizeof(DevResult), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->h_E, h->d_E, sizeof(float) * h->D * h->M, cudaMemcpyDeviceToHost, h->stream));
    return enqueue_fetch_windows(h);
}
static int enqueue_fetch_windows(pcs_handle* h) {
    CUDA_TRY(cudaMemcpyAsync(h->h_sigwin, h->d_sigwin, sizeof(float2) * h->win_cap, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->h_noisewin, h->d_noisewin, sizeof(float2) * h->win_cap, cudaMemcpyDeviceToHost, h->stream));
    return 0;
}
static int enqueue_fetch_demod(pcs_handle* h) {
    CUDA_TRY(cudaMemcpyAsync(h->h_res, h->d_res, sizeof(DevResult), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->h_sym, h->d_sym, sizeof(int) * h->max_sym, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->h_centre, h->d_centre, sizeof(int) * h->max_sym, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->h_mag, h->d_mag, sizeof(float) * h->max_sym, cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

static void copy_result(const pcs_handle* h, pcs_result* res) {
    static_assert(sizeof(pcs_result) == sizeof(DevResult), "pcs_result and DevResult must have the same layout");
    if (res) memcpy(res, h->h_res, sizeof(pcs_result));
}
static void copy_search_out(const pcs_handle* h, pcs_result* res, float* E_out) {
    copy_result(h, res);
    if (E_out) memcpy(E_out, h->h_E, sizeof(float) * h->D * h->M);
}
static void copy_demod_out(const pcs_handle* h, pcs_result* res, int32_t* sym, int32_t* centre, float* mag) {
    copy_result(h, res);
    const int n = std::min(std::max(h->h_res->n_sym, 0), h->max_sym);
    if (sym) memcpy(sym, h->h_sym, sizeof(int) * n);
    if (centre) memcpy(centre, h->h_centre, sizeof(int) * n);
    if (mag) memcpy(mag, h->h_mag, sizeof(float) * n);
}

// ---- whole-chunk CUDA graph ------------------------------------------------------------------------------
static bool graph_usable(const pcs_handle* h) {
    return h->graph_enabled && !h->graph_failed && !h->profiling && h->bin_lo == 0 && h->bin_hi == h->D &&
           h->d_x_cur == h->d_x && h->eager_chunks >= 1;
}

static int enqueue_chunk_eager(pcs_handle* h) {
    if (int rc = enqueue_search(h)) return rc;
    if (int rc = enqueue_demod(h, -1, false)) return rc;
    if (int rc = enqueue_fetch_search(h)) return rc;
    return enqueue_fetch_demod(h);
}

// The same work as enqueue_chunk_eager shaped as a fork-join for capture: the chunk spectrum (pass 1) and, once the
// estimate is known, the SNR bins and their D2H run on a side branch; the main branch goes search -> estimate ->
// demod -> timing -> symbols -> D2H without waiting for them.
static int enqueue_chunk_forked(pcs_handle* h) {
    cudaStream_t main_s = h->stream;
    CUDA_TRY(cudaEventRecord(h->ev_fork, main_s));
    CUDA_TRY(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    int rc = 0;
    if (!h->parseval) {            // (the Parseval variant computes the full spectrum on the main branch anyway)
        h->stream = h->side;
        rc = enqueue_spectrum(h);
        h->stream = main_s;
        if (rc) return rc;
    }
    if ((rc = enqueue_search_local(h))) return rc;
    if ((rc = enqueue_estimate_kernel(h))) return rc;
    CUDA_TRY(cudaEventRecord(h->ev_est, main_s));
    CUDA_TRY(cudaStreamWaitEvent(h->side, h->ev_est, 0));
    h->stream = h->side;
    rc = enqueue_snr_bins(h);
    if (!rc) rc = enqueue_fetch_windows(h);
    h->stream = main_s;
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->ev_side, h->side));
    if ((rc = enqueue_demod(h, -1, false))) return rc;
    CUDA_TRY(cudaMemcpyAsync(h->h_E, h->d_E, sizeof(float) * h->D * h->M, cudaMemcpyDeviceToHost, main_s));
    if ((rc = enqueue_fetch_demod(h))) return rc;
    CUDA_TRY(cudaStreamWaitEvent(main_s, h->ev_side, 0));
    return 0;
}

// search + estimate + demod + D2H of the results for the chunk in d_x, as one graph launch when possible.
static int enqueue_chunk(pcs_handle* h) {
    if (graph_usable(h) && !h->gexec) {
        cudaGraph_t graph = nullptr;
        const int64_t before = h->launches;
        if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            const int rc = enqueue_chunk_forked(h);
            const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
            if (rc == 0 && ce == cudaSuccess && graph &&
                cudaGraphInstantiate(&h->gexec, graph, 0) == cudaSuccess) {
                h->graph_launches = h->launches - before;
            } else {
                h->gexec = nullptr;
                h->graph_failed = true;
                cudaGetLastError();
            }
            if (graph) cudaGraphDestroy(graph);
            h->launches = before;
        } else {
            h->graph_failed = true;
            cudaGetLastError();
        }
    }
    if (graph_usable(h) && h->gexec) {
        CUDA_TRY(cudaGraphLaunch(h->gexec, h->stream));
        h->launches += h->graph_launches;
        h->spectrum_pending = false;
    } else {
        if (int rc = enqueue_chunk_eager(h)) return rc;
        h->eager_chunks++;
    }
    h->fetch_in_flight = true;
    return 0;
}

int pcs_upload(pcs_handle* h) {
    if (!h) return fail(PCS_ERR_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const float2* from = h->h_src ? h->h_src : h->h_x;
    h->h_src = nullptr;
    CUDA_TRY(cudaMemcpyAsync(h->d_x, from, sizeof(float2) * h->N, cudaMemcpyHostToDevice, h->stream));
    h->d_x_cur = h->d_x_base = h->d_x;
    h->uploaded = true;
    h->searched = h->demodulated = false;
    h->spectrum_pending = true;        // enqueued with the estimate (eager) or on the graph's side branch
    h->spectrum_full = false;
    return PCS_OK;
}

// a19 + a4 for the STX backend (STX.py:13-20, dem_base:670-707): the chunk goes to HBM, is clipped there (five small
// kernels, thresholds by NumPy's own summation order), and comes back into the pinned buffer because the caller carries
// raw[-overlap:] into the next chunk (demodulator_process.py:337) and the reference clips in place.
int pcs_upload_thresholded(pcs_handle* h, float scale, int64_t* clipped_idx, int32_t cap, int32_t* n_clipped,
                           float* thresholds) {
    if (!h || !n_clipped || (cap > 0 && !clipped_idx) || cap < 0) return fail(PCS_ERR_INVALID, "bad argument");
    const int N = h->N, nb = N / 1024;
    if (nb > 4096) return fail(PCS_ERR_INVALID, "pcs_upload_thresholded supports nfft <= 2^22");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (!h->d_thr_partial) {
        if (int rc = dev_alloc(h, &h->d_thr_partial, (size_t)nb)) return rc;
        if (int rc = dev_alloc(h, &h->d_thr_level, (size_t)2)) return rc;
        if (int rc = dev_alloc(h, &h->d_thr_bits, (size_t)N / 32)) return rc;
        CUDA_TRY(cudaHostAlloc((void**)&h->h_thr_bits, sizeof(unsigned int) * (N / 32) + 2 * sizeof(float), cudaHostAllocDefault));
    }
    h->h_src = nullptr;       // clipping is in place in the handle's own buffer: a pending caller source does not apply
    float* mag = h->d_p;      // free until the demod stage of this chunk writes it
    CUDA_TRY(cudaMemcpyAsync(h->d_x, h->h_x, sizeof(float2) * N, cudaMemcpyHostToDevice, h->stream));
    threshold_abs_kernel<<<nb, 128, 0, h->stream>>>(h->d_x, mag, h->d_thr_partial);
    threshold_level_kernel<<<1, 1024, 0, h->stream>>>(h->d_thr_partial, nb, scale, 1.0f / (float)N, h->d_thr_level);
    threshold_clip_kernel<false><<<nb, 128, 0, h->stream>>>(h->d_x, mag, h->d_thr_level, h->d_thr_partial, nullptr);
    threshold_level_kernel<<<1, 1024, 0, h->stream>>>(h->d_thr_partial, nb, scale, 1.0f / (float)N, h->d_thr_level + 1);
    threshold_clip_kernel<true><<<nb, 128, 0, h->stream>>>(h->d_x, mag, h->d_thr_level + 1, nullptr, h->d_thr_bits);
    h->launches += 5;
    CUDA_TRY(cudaGetLastError());
    float* h_levels = reinterpret_cast<float*>(h->h_thr_bits + N / 32);
    CUDA_TRY(cudaMemcpyAsync(h->h_thr_bits, h->d_thr_bits, sizeof(unsigned int) * (N / 32), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h_levels, h->d_thr_level, 2 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->h_x, h->d_x, sizeof(float2) * N, cudaMemcpyDeviceToHost, h->stream));
    h->d_x_cur = h->d_x_base = h->d_x;
    h->uploaded = true;
    h->searched = h->demodulated = false;
    h->spectrum_pending = true;
    h->spectrum_full = false;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    int n = 0;
    for (int w = 0; w < N / 32; ++w) {
        unsigned int word = h->h_thr_bits[w];
        while (word) {
            const int b = __builtin_ctz(word);
            word &= word - 1;
            if (n < cap) clipped_idx[n] = (int64_t)w * 32 + b;
            ++n;
        }
    }
    *n_clipped = n;       // may exceed cap: the caller then knows the list was truncated
    if (thresholds) { thresholds[0] = h_levels[0]; thresholds[1] = h_levels[1]; }
    return PCS_OK;
}

int pcs_upload_device(pcs_handle* h, const void* d_chunk) {
    if (!h || !d_chunk) return fail(PCS_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    h->d_x_cur = h->d_x_base = reinterpret_cast<const float2*>(d_chunk);
    h->uploaded = true;
    h->searched = h->demodulated = false;
    h->spectrum_pending = true;
    h->spectrum_full = false;
    return PCS_OK;
}

// Doppler-rate hypothesis (extension; kern:755-778): the chunk as uploaded, times exp(j (a n^2 + b n + c)), becomes the chunk
// every following pcs_search / pcs_demod / pcs_process works on, until the next upload or pcs_heterodyne call (each call
// starts from the chunk as uploaded; a = b = c = 0 restores it without a copy).
int pcs_heterodyne(pcs_handle* h, float a, float b, float c) {
    if (!h) return fail(PCS_ERR_INVALID, "null handle");
    if (!h->uploaded || !h->d_x_base) return fail(PCS_ERR_STATE, "pcs_heterodyne before pcs_upload");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    h->searched = h->demodulated = false;
    h->spectrum_pending = true;
    h->spectrum_full = false;
    if (a == 0.f && b == 0.f && c == 0.f) {
        h->d_x_cur = h->d_x_base;
        return PCS_OK;
    }
    if (!h->d_xh)
        if (int rc = dev_alloc(h, &h->d_xh, (size_t)h->N)) return rc;
    heterodyne_kernel<<<(h->N + 255) / 256, 256, 0, h->stream>>>(h->d_x_base, h->d_xh, a, b, c, h->N);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    h->d_x_cur = h->d_xh;
    return PCS_OK;
}

// The chunk the search currently works on (complex64[nfft]): inspection hook for pcs_heterodyne.
int pcs_get_chunk(pcs_handle* h, float* x_out) {
    if (!h || !x_out) return fail(PCS_ERR_INVALID, "null argument");
    if (!h->uploaded) return fail(PCS_ERR_STATE, "no chunk uploaded");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaMemcpyAsync(x_out, h->d_x_cur, sizeof(float2) * h->N, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PCS_OK;
}

int pcs_search(pcs_handle* h, pcs_result* res, float* E_out) {
    if (!h) return fail(PCS_ERR_INVALID, "null handle");
    if (!h->uploaded) return fail(PCS_ERR_STATE, "pcs_search before pcs_upload");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (int rc = enqueue_search(h)) return rc;
    if (int rc = enqueue_fetch_search(h)) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->searched = true;
    copy_search_out(h, res, E_out);
    return PCS_OK;
}

int pcs_demod(pcs_handle* h, int32_t shift, pcs_result* res, int32_t* sym, int32_t* centre, float* mag) {
    if (!h) return fail(PCS_ERR_INVALID, "null handle");
    if (!h->uploaded) return fail(PCS_ERR_STATE, "pcs_demod before pcs_upload");
    if (shift < 0 && !h->searched) return fail(PCS_ERR_STATE, "pcs_demod(shift<0) needs a preceding pcs_search");
    if (shift >= h->N) return fail(PCS_ERR_INVALID, "shift %d outside [0, nfft)", shift);
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (int rc = enqueue_demod(h, shift, false)) return rc;
    if (int rc = enqueue_fetch_demod(h)) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->demodulated = true;
    h->h_res->demod_shift = shift >= 0 ? shift : h->h_res->shift;
    copy_demod_out(h, res, sym, centre, mag);
    return PCS_OK;
}

int pcs_process(pcs_handle* h, pcs_result* res, float* E_out, int32_t* sym, int32_t* centre, float* mag) {
    if (!h) return fail(PCS_ERR_INVALID, "null handle");
    if (!h->uploaded) return fail(PCS_ERR_STATE, "pcs_process before pcs_upload");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (int rc = enqueue_chunk(h)) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->fetch_in_flight = false;
    h->searched = h->demodulated = true;
    h->h_res->demod_shift = h->h_res->shift;
    copy_search_out(h, res, E_out);
    copy_demod_out(h, res, sym, centre, mag);
    return PCS_OK;
}

int pcs_enqueue_device(pcs_handle* h, const void* d_chunk) {
    if (!h || !d_chunk) return fail(PCS_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (h->graph_enabled && !h->graph_failed && !h->profiling && h->bin_lo == 0 && h->bin_hi == h->D) {
        // graph mode: the captured kernels read the handle's own chunk buffer, so stage the chunk there (one 8N-byte
        // device-to-device copy, ~1 us per MB)
        CUDA_TRY(cudaMemcpyAsync(h->d_x, d_chunk, sizeof(float2) * h->N, cudaMemcpyDeviceToDevice, h->stream));
        h->d_x_cur = h->d_x_base = h->d_x;
        h->uploaded = true;
        h->spectrum_pending = true;
        h->spectrum_full = false;
        if (int rc = enqueue_chunk(h)) return rc;
    } else {
        if (int rc = pcs_upload_device(h, d_chunk)) return rc;
        if (int rc = enqueue_search(h)) return rc;
        if (int rc = enqueue_demod(h, -1, false)) return rc;
        h->fetch_in_flight = false;
    }
    h->searched = h->demodulated = true;
    return PCS_OK;
}

int pcs_fetch(pcs_handle* h, pcs_result* res, float* E_out, int32_t* sym, int32_t* centre, float* mag) {
    if (!h) return fail(PCS_ERR_INVALID, "null handle");
    if (!h->demodulated && !h->fetch_in_flight) return fail(PCS_ERR_STATE, "pcs_fetch before a chunk was enqueued");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (!h->fetch_in_flight) {
        if (int rc = enqueue_fetch_search(h)) return rc;
        if (int rc = enqueue_fetch_demod(h)) return rc;
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->fetch_in_flight = false;
    h->h_res->demod_shift = h->h_res->shift;
    copy_search_out(h, res, E_out);
    copy_demod_out(h, res, sym, centre, mag);
    return PCS_OK;
}

int pcs_snr_windows(pcs_handle* h, float* sig_win, float* noise_win) {
    if (!h) return fail(PCS_ERR_INVALID, "null handle");
    if (!h->searched) return fail(PCS_ERR_STATE, "pcs_snr_windows before pcs_search");
    if (sig_win) memcpy(sig_win, h->h_sigwin, sizeof(float2) * std::max(h->h_res->sig_len, 0));
    if (noise_win) memcpy(noise_win, h->h_noisewin, sizeof(float2) * std::max(h->h_res->noise_len, 0));
    return PCS_OK;
}

// mean(|z|) of a complex64 window as float32: magnitudes by hypotf, accumulated in double, rounded once.  (NumPy's own
// np.abs / np.mean differ from this in the last ulp depending on the host's SIMD dispatch; every schedule of this
// library -- two-step, fused, one-call, streaming -- goes through this one function, so they agree exactly.)
static float mean_abs(const float2* z, int n) {
    double a = 0;
    for (int i = 0; i < n; ++i) a += (double)hypotf(z[i].x, z[i].y);
    return (float)(a / n);
}

// computeSNR's two window means (dem_base:657-663) for the common geometry in which neither window touches the ends
// of the spectrum, so that the reference's slices X[a-w : b+w] are exactly the gathered windows.  *ok = 0 otherwise
// (the caller then evaluates the reference's slicing rules itself).
static void snr_means_from(int N, int w, const int32_t* shifts, const pcs_result* r, const float2* sigwin,
                           const float2* noisewin, float* sig_mean, float* noise_mean, int32_t* ok) {
    *ok = 0;
    if (r->status != 0 || r->sig_len <= 0) return;
    const int lo = shifts[r->low_idx], hi = shifts[r->high_idx];
    const int nlo = (lo + N / 2) % N, nhi = (hi + N / 2) % N;
    if (!(lo <= hi && nlo <= nhi && lo - w >= 0 && nlo - w >= 0 && hi + w <= N && nhi + w <= N && r->sig_start == lo - w &&
          r->noise_start == nlo - w && r->sig_len == hi - lo + 2 * w && r->noise_len == nhi - nlo + 2 * w))
        return;
    *sig_mean = mean_abs(sigwin, r->sig_len);
    *noise_mean = mean_abs(noisewin, r->noise_len);
    *ok = 1;
}

int pcs_snr_means(pcs_handle* h, const int32_t* shifts, float* sig_mean, float* noise_mean, int32_t* ok) {
    if (!h || !shifts || !sig_mean || !noise_mean || !ok) return fail(PCS_ERR_INVALID, "null argument");
    if (!h->searched) return fail(PCS_ERR_STATE, "pcs_snr_means before a search");
    snr_means_from(h->N, h->cfg.snr_window, shifts, reinterpret_cast<const pcs_result*>(h->h_res), h->h_sigwin, h->h_noisewin,
                   sig_mean, noise_mean, ok);
    return PCS_OK;
}

int pcs_mean_abs_c64(const float* z, int32_t n, float* out) {
    if (!z || !out || n < 1) return fail(PCS_ERR_INVALID, "pcs_mean_abs_c64: bad argument");
    *out = mean_abs(reinterpret_cast<const float2*>(z), n);
    return PCS_OK;
}

// One call per chunk for a host that wants bits: pcs_upload + pcs_process + pcs_snr_means + pcs_stitch_chunk, with the
// symbol tables handed from the result staging area to the stitcher without leaving the library.  sym / centre / mag
// (inspection copies) and E_out may be NULL.
int pcs_chunk_to_bits(pcs_handle* h, pcs_stitcher* st, const int32_t* shifts, const int64_t* clipped, int32_t n_clipped,
                      pcs_result* res, float* E_out, int32_t* sym, int32_t* centre, float* mag, float* sig_mean,
                      float* noise_mean, int32_t* snr_ok, uint8_t* bits_out, uint8_t* centres_out, uint8_t* trust_out,
                      int32_t* n_out) {
    if (!h || !st || !res || !n_out) return fail(PCS_ERR_INVALID, "null argument");
    if (int rc = pcs_upload(h)) return rc;
    if (int rc = pcs_process(h, res, E_out, sym, centre, mag)) return rc;
    *n_out = 0;
    if (snr_ok) {
        if (int rc = pcs_snr_means(h, shifts, sig_mean, noise_mean, snr_ok)) return rc;
    }
    const int n = std::min(std::max(h->h_res->n_sym, 0), h->max_sym);
    *n_out = -1;      // from here on a failure is the stitcher's: res, E and the symbol tables are valid
    return pcs_stitch_chunk(st, h->h_sym, h->h_centre, h->h_mag, n, clipped, n_clipped, h->h_res->sp_sym, bits_out, centres_out,
                            trust_out, n_out);
}

int pcs_get_spectrum(pcs_handle* h, float* X_out) {
    if (!h || !X_out) return fail(PCS_ERR_INVALID, "null argument");
    if (!h->uploaded) return fail(PCS_ERR_STATE, "no chunk uploaded");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (!h->spectrum_full) {
        if (int rc = fft_large<-1>(h, LoadC{h->d_x_cur}, h->d_X)) return rc;
        h->spectrum_full = true;
    }
    CUDA_TRY(cudaMemcpyAsync(X_out, h->d_X, sizeof(float2) * h->N, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PCS_OK;
}

int pcs_get_peaks(pcs_handle* h, float* peak_val, int32_t* peak_offset) {
    if (!h) return fail(PCS_ERR_INVALID, "null handle");
    if (!h->searched) return fail(PCS_ERR_STATE, "pcs_get_peaks before pcs_search");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const size_t n = (size_t)h->D * h->M;
    if (peak_val) CUDA_TRY(cudaMemcpyAsync(peak_val, h->d_peakv, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
    if (peak_offset) CUDA_TRY(cudaMemcpyAsync(peak_offset, h->d_peako, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PCS_OK;
}

int pcs_get_demod_surface(pcs_handle* h, int32_t shift, float* y_out) {
    if (!h || !y_out) return fail(PCS_ERR_INVALID, "null argument");
    if (!h->uploaded) return fail(PCS_ERR_STATE, "no chunk uploaded");
    if (shift < 0 || shift >= h->N) return fail(PCS_ERR_INVALID, "shift %d outside [0, nfft)", shift);
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (!h->d_ycplx)
        if (int rc = dev_alloc(h, &h->d_ycplx, (size_t)h->M * h->N)) return rc;
    if (int rc = enqueue_demod(h, shift, true)) return rc;
    CUDA_TRY(cudaMemcpyAsync(y_out, h->d_ycplx, sizeof(float2) * (size_t)h->M * h->N, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PCS_OK;
}

int pcs_get_demod_magnitudes(pcs_handle* h, float* ymag_out, float* p_out) {
    if (!h) return fail(PCS_ERR_INVALID, "null handle");
    if (!h->demodulated) return fail(PCS_ERR_STATE, "no chunk demodulated");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (ymag_out)
        CUDA_TRY(cudaMemcpyAsync(ymag_out, h->d_ymag, sizeof(float) * (size_t)h->M * h->N, cudaMemcpyDeviceToHost, h->stream));
    if (p_out) CUDA_TRY(cudaMemcpyAsync(p_out, h->d_p, sizeof(float) * h->N, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PCS_OK;
}

int pcs_set_stream(pcs_handle* h, uint64_t stream) {
    if (!h) return fail(PCS_ERR_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (h->own_stream) CUDA_TRY(cudaStreamDestroy(h->stream));
    h->stream = reinterpret_cast<cudaStream_t>((uintptr_t)stream);
    h->own_stream = false;
    return PCS_OK;
}

int pcs_set_profiling(pcs_handle* h, int enable) {
    if (!h) return fail(PCS_ERR_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (enable && !h->ev[0])
        for (cudaEvent_t& e : h->ev) CUDA_TRY(cudaEventCreate(&e));
    if (!enable) CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->profiling = enable != 0;
    for (int s = 0; s < PCS_NUM_STAGES; ++s) {
        h->ev_pending[s] = false;
        h->stage_ms[s] = 0;
        h->stage_count[s] = 0;
    }
    return PCS_OK;
}

int pcs_get_profile(pcs_handle* h, double* stage_ms, int64_t* stage_count) {
    if (!h) return fail(PCS_ERR_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    for (int s = 0; s < PCS_NUM_STAGES; ++s) {
        if (h->ev_pending[s]) {
            float ms = 0.f;
            CUDA_TRY(cudaEventSynchronize(h->ev[2 * s + 1]));
            CUDA_TRY(cudaEventElapsedTime(&ms, h->ev[2 * s], h->ev[2 * s + 1]));
            h->stage_ms[s] += ms;
            h->stage_count[s]++;
            h->ev_pending[s] = false;
        }
        if (stage_ms) stage_ms[s] = h->stage_ms[s];
        if (stage_count) stage_count[s] = h->stage_count[s];
    }
    return PCS_OK;
}

// fp32 FMA throughput of the device (roofline denominator for the FFT-bound kernels).
int pcs_measure_fp32_peak(int device, double* tflops) {
    if (!tflops) return fail(PCS_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    float* sink = nullptr;
    CUDA_TRY(cudaMalloc(&sink, sizeof(float) * 4096));
    const int grid = prop.multiProcessorCount * 8, block = 256, iters = 1 << 14;
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(a));
        fma_peak_kernel<<<grid, block>>>(sink, iters, 1.0001f, 0.9999f);
        CUDA_TRY(cudaEventRecord(b));
        CUDA_TRY(cudaEventSynchronize(b));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
        const double flops = 2.0 * 16 * (double)iters * grid * block;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(sink);
    *tflops = best;
    return PCS_OK;
}

// ---- bin sharding (one process per GPU; the exchange itself is an NCCL all-gather issued by the caller) ----
int pcs_set_bin_range(pcs_handle* h, int32_t lo, int32_t hi) {
    if (!h) return fail(PCS_ERR_INVALID, "null handle");
    if (lo < 0 || hi > h->D || lo >= hi) return fail(PCS_ERR_INVALID, "bin range [%d, %d) outside [0, %d)", lo, hi, h->D);
    h->bin_lo = lo;
    h->bin_hi = hi;
    return PCS_OK;
}

int pcs_shard_buffers(pcs_handle* h, void** d_energy, void** d_peak_val, void** d_peak_off) {
    if (!h) return fail(PCS_ERR_INVALID, "null handle");
    if (d_energy) *d_energy = h->d_Efull;
    if (d_peak_val) *d_peak_val = h->d_peakv;
    if (d_peak_off) *d_peak_off = h->d_peako;
    return PCS_OK;
}

int pcs_enqueue_search_local(pcs_handle* h) {
    if (!h) return fail(PCS_ERR_INVALID, "null handle");
    if (!h->uploaded) return fail(PCS_ERR_STATE, "search before upload");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    return enqueue_search_local(h);
}

int pcs_enqueue_estimate_and_demod(pcs_handle* h, int32_t with_demod) {
    if (!h) return fail(PCS_ERR_INVALID, "null handle");
    if (!h->uploaded) return fail(PCS_ERR_STATE, "estimate before upload");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (int rc = enqueue_estimate(h)) return rc;
    h->searched = true;
    if (with_demod) {
        if (int rc = enqueue_demod(h, -1, false)) return rc;
        h->demodulated = true;
    }
    return PCS_OK;
}

#include "shard.inc"

}  // extern "C"
