// COMPARISON VARIANT ONLY (BASELINE.json north_star: "a cuFFT+LTO-callback build serves only as an internal comparison
// variant"; SURVEY.md 2.2(ii) bar 2, 7 step 8): rows a6-a8 of SURVEY 8(a) as a batched inverse cuFFT with load / store
// callbacks, so that the two elementwise kernels of the reference disappear into the transform:
//   load  callback  = multInputVectorWithShiftedMasksDopp (kern:339-373): Y[d,m,k] = X[(k + s_d) % N] * Mk[m,k], computed
//                     on the fly instead of being written to HBM first;
//   store callback  = blockAbsSumAtomic (kern:421-480): |y|^2 / 2^18 accumulated (float atomics, like the reference) into
//                     1024 partial sums per transform instead of re-reading the surface.
// Callback flavour: cuFFT's classic device-pointer callbacks (cufftXtSetCallback; static libcufft + relocatable device
// code).  The LTO flavour (cufftXtSetJITCallback + nvJitLink) was tried first: on this image (cuFFT 11.4.1, B200)
// cufftMakePlan1d returns CUFFT_INTERNAL_ERROR for every plan that carries a JIT callback, for lto_100 and lto_100a IR alike
// (tools/ubench/lto_probe/).  Builds into libpcs_cufft_variant.so (~ 290 MB: the static cuFFT), separate from the product
// library, ON THE GPU BOX (`make -C pycusdr_b200/csrc variant`; tools/cufft_variant.py does it), and is reported under
// `variants` only.  Dataflow per batch of Doppler bins: cuFFT's first pass pulls the shifted-spectrum x filter product
// through the load callback, its intermediate passes go through HBM (in place, the batch's [bins x M x N] buffer), its last
// pass hands every output sample to the store callback.  Unlike the product's fused overlap-save kernel the surface still
// crosses HBM between cuFFT's passes.
#include <cuda_runtime.h>
#include <cufft.h>
#include <cufftXt.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

struct PcsvInfo {
    const float2* X;        // [N] chunk spectrum
    const float2* Mk;       // [M][N] conjugate filter spectra
    const int* shifts;      // [D]
    float* Epart;           // [batch transforms][1024] partial energies
    unsigned int nmask;     // N - 1
    int logN, M, bin0;      // first Doppler bin of the batch
};

__device__ cufftComplex pcsv_load(void* dataIn, size_t offset, void* callerInfo, void* sharedPointer) {
    const PcsvInfo* p = static_cast<const PcsvInfo*>(callerInfo);
    const unsigned int k = (unsigned int)offset & p->nmask;
    const unsigned int b = (unsigned int)(offset >> p->logN);
    const int d = p->bin0 + (int)(b / (unsigned int)p->M), m = (int)(b % (unsigned int)p->M);
    const float2 x = p->X[(k + (unsigned int)p->shifts[d]) & p->nmask];
    const float2 g = p->Mk[((size_t)m << p->logN) | k];
    return make_float2(x.x * g.x - x.y * g.y, x.x * g.y + x.y * g.x);
}

__device__ void pcsv_store(void* dataOut, size_t offset, cufftComplex v, void* callerInfo, void* sharedPointer) {
    const PcsvInfo* p = static_cast<const PcsvInfo*>(callerInfo);
    const unsigned int b = (unsigned int)(offset >> p->logN);
    atomicAdd(&p->Epart[((size_t)b << 10) | ((unsigned int)offset & 1023u)], (v.x * v.x + v.y * v.y) * (1.0f / 262144.0f));
}

__device__ cufftCallbackLoadC d_pcsv_load = pcsv_load;
__device__ cufftCallbackStoreC d_pcsv_store = pcsv_store;

static std::string g_err;
#define CK(expr)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) { g_err = std::string(#expr) + ": " + cudaGetErrorString(e__); return -2; } \
    } while (0)
#define CKF(expr)                                                                                 \
    do {                                                                                          \
        cufftResult r__ = (expr);                                                                 \
        if (r__ != CUFFT_SUCCESS) { g_err = std::string(#expr) + ": cufftResult " + std::to_string((int)r__); return -3; } \
    } while (0)

__global__ void pcsv_reduce_kernel(const float* __restrict__ Epart, float* __restrict__ E, int rows) {
    // one warp per transform: 1024 partials in a fixed order
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= rows) return;
    float s = 0.f;
    for (int i = lane; i < 1024; i += 32) s += Epart[(size_t)w * 1024 + i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) E[w] = s;
}

extern "C" const char* pcsv_last_error(void) { return g_err.c_str(); }

// x: complex64[N] chunk (host), masks: complex64[M][N] (host), shifts: int32[D].  Runs the search `reps` times (after one
// warm-up) and returns E[D][M] (per-mask energies, kern:442 scale) of the last run and the mean device time per run.
extern "C" int pcsv_search(const float* x, int N, const float* masks, int M, const int32_t* shifts, int D, int bins_per_batch,
                           int reps, float* E_out, float* ms_per_search) {
    if (!x || !masks || !shifts || !E_out || N < 2 || (N & (N - 1)) || M < 1 || D < 1) { g_err = "bad argument"; return -1; }
    int logN = 0;
    while ((1 << logN) < N) ++logN;
    if (bins_per_batch < 1 || bins_per_batch > D) bins_per_batch = D;
    while (D % bins_per_batch) --bins_per_batch;
    const int batch = bins_per_batch * M;
    float2 *d_x = nullptr, *d_X = nullptr, *d_M = nullptr, *d_buf = nullptr;
    int* d_s = nullptr;
    float *d_Epart = nullptr, *d_E = nullptr;
    PcsvInfo* d_info = nullptr;
    CK(cudaMalloc(&d_x, sizeof(float2) * N));
    CK(cudaMalloc(&d_X, sizeof(float2) * N));
    CK(cudaMalloc(&d_M, sizeof(float2) * (size_t)M * N));
    CK(cudaMalloc(&d_buf, sizeof(float2) * (size_t)batch * N));
    CK(cudaMalloc(&d_s, sizeof(int) * D));
    CK(cudaMalloc(&d_Epart, sizeof(float) * (size_t)batch * 1024));
    CK(cudaMalloc(&d_E, sizeof(float) * (size_t)D * M));
    CK(cudaMalloc(&d_info, sizeof(PcsvInfo)));
    CK(cudaMemcpy(d_x, x, sizeof(float2) * N, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_M, masks, sizeof(float2) * (size_t)M * N, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_s, shifts, sizeof(int) * D, cudaMemcpyHostToDevice));
    cufftHandle fwd, inv;
    CKF(cufftPlan1d(&fwd, N, CUFFT_C2C, 1));
    CKF(cufftCreate(&inv));
    size_t ws = 0;
    CKF(cufftMakePlan1d(inv, N, CUFFT_C2C, batch, &ws));
    void* info_ptr = d_info;
    cufftCallbackLoadC h_ld;
    cufftCallbackStoreC h_st;
    CK(cudaMemcpyFromSymbol(&h_ld, d_pcsv_load, sizeof(h_ld)));
    CK(cudaMemcpyFromSymbol(&h_st, d_pcsv_store, sizeof(h_st)));
    CKF(cufftXtSetCallback(inv, (void**)&h_ld, CUFFT_CB_LD_COMPLEX, &info_ptr));
    CKF(cufftXtSetCallback(inv, (void**)&h_st, CUFFT_CB_ST_COMPLEX, &info_ptr));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float total_ms = 0.f;
    for (int rep = 0; rep <= reps; ++rep) {
        if (rep == 1) CK(cudaEventRecord(e0));
        CKF(cufftExecC2C(fwd, d_x, d_X, CUFFT_FORWARD));
        for (int bin0 = 0; bin0 < D; bin0 += bins_per_batch) {
            PcsvInfo info{d_X, d_M, d_s, d_Epart, (unsigned int)N - 1u, logN, M, bin0};
            CK(cudaMemcpyAsync(d_info, &info, sizeof(info), cudaMemcpyHostToDevice));
            CK(cudaMemsetAsync(d_Epart, 0, sizeof(float) * (size_t)batch * 1024));
            CKF(cufftExecC2C(inv, d_buf, d_buf, CUFFT_INVERSE));
            pcsv_reduce_kernel<<<(batch * 32 + 255) / 256, 256>>>(d_Epart, d_E + (size_t)bin0 * M, batch);
            CK(cudaGetLastError());
        }
    }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&total_ms, e0, e1));
    if (ms_per_search) *ms_per_search = reps > 0 ? total_ms / reps : 0.f;
    CK(cudaMemcpy(E_out, d_E, sizeof(float) * (size_t)D * M, cudaMemcpyDeviceToHost));
    cufftDestroy(fwd);
    cufftDestroy(inv);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    for (void* p : {(void*)d_x, (void*)d_X, (void*)d_M, (void*)d_buf, (void*)d_s, (void*)d_Epart, (void*)d_E, (void*)d_info}) cudaFree(p);
    return 0;
}
