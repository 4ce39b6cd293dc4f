// Micro-benchmark: throughput of scalar vs packed (f32x2) fp32 instructions on sm_100a, and how they co-issue with
// ALU-pipe work.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2 fp32x2.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
template <int MODE>
__global__ void k(float* out, float a, float b) {
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    const float2 A = make_float2(a, a), B = make_float2(b, b);
    int acc = threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { v[i].x = fmaf(v[i].x, a, b); v[i].y = fmaf(v[i].y, a, b); }          // 2 FFMA
            if (MODE == 1) { v[i] = __ffma2_rn(v[i], A, B); }                                      // 1 FFMA2
            if (MODE == 2) { v[i].x = v[i].x + b; v[i].y = v[i].y + b; }                           // 2 FADD
            if (MODE == 3) { v[i] = __fadd2_rn(v[i], B); }                                         // 1 FADD2
            if (MODE == 4) { v[i].x = v[i].x * a; v[i].y = v[i].y * a; }                           // 2 FMUL
            if (MODE == 5) { v[i] = __fmul2_rn(v[i], A); }                                         // 1 FMUL2
            if (MODE == 6) { v[i] = __fadd2_rn(v[i], B); acc = (acc ^ (acc << 1)) + i; }           // FADD2 + ALU
            if (MODE == 7) { v[i].x = v[i].x + b; v[i].y = v[i].y + b; acc = (acc ^ (acc << 1)) + i; }
            if (MODE == 8) { v[i] = __fadd2_rn(v[i], B); v[(i + 4) & 7].x = fmaf(v[(i + 4) & 7].x, a, b); }  // FADD2 + FFMA
        }
    }
    float s = acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i].x + v[i].y;
    if (s == 123.456f) out[threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double flop_per_inner) {
    float* d; cudaMalloc(&d, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, block = 256;
    k<MODE><<<grid, block>>>(d, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<grid, block>>>(d, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inner = (double)ITERS * 8 * grid * block;
    printf("%-28s %8.3f ms  %7.2f Ginner/s  %7.2f TFLOP/s  (%.3f inner/clk/SM at 1.9 GHz)\n", name, ms, inner / ms / 1e6,
           inner * flop_per_inner / ms / 1e9, inner / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(d);
}

int main() {
    run<0>("2xFFMA", 4); run<1>("FFMA2", 4); run<2>("2xFADD", 2); run<3>("FADD2", 2); run<4>("2xFMUL", 2); run<5>("FMUL2", 2);
    run<6>("FADD2+2 ALU", 2); run<7>("2xFADD+2 ALU", 2); run<8>("FADD2+FFMA", 4);
    return 0;
}
