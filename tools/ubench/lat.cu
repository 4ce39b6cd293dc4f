// Micro-benchmark: issue rate of FADD2 / FFMA2 / FFMA as a function of warps per SM sub-partition and independent
// chains per thread (latency / parallelism needed to keep the FMA pipe busy).  sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE, int ILP>
__global__ void k(float* out, float a, float b) {
    float2 v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    const float2 A = make_float2(a, a), B = make_float2(b, b);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep)
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (MODE == 0) v[i].x = fmaf(v[i].x, a, b);
                if (MODE == 1) v[i] = __ffma2_rn(v[i], A, B);
                if (MODE == 2) v[i] = __fadd2_rn(v[i], B);
                if (MODE == 3) v[i].x = v[i].x + b;
            }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i].x + v[i].y;
    if (s == 123.456f) out[threadIdx.x] = s;
}
template <int MODE, int ILP>
void run(const char* name, int warps_per_smsp) {
    float* d; cudaMalloc(&d, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int block = 32 * 4 * warps_per_smsp, grid = 148;      // one CTA per SM
    k<MODE, ILP><<<grid, block>>>(d, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<MODE, ILP><<<grid, block>>>(d, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = (double)ITERS * 8 * ILP * grid * (block / 32);
    printf("%-6s ilp %d warps/smsp %2d : %.3f warp-inst/clk/smsp\n", name, ILP, warps_per_smsp, inst / (ms * 1e-3) / 148 / 4 / 1.92e9);
    cudaFree(d);
}
template <int MODE> void sweep(const char* n) {
    for (int w : {1, 2, 4, 8}) { run<MODE, 1>(n, w); run<MODE, 2>(n, w); run<MODE, 4>(n, w); run<MODE, 8>(n, w); }
}
int main() { sweep<0>("FFMA"); sweep<3>("FADD"); sweep<1>("FFMA2"); sweep<2>("FADD2"); return 0; }
