// Egress bandwidth of a kernel that broadcasts a buffer to peer GPUs with plain stores (the engine's chunk_bcast_kernel):
// grid size x unroll x number of destinations, plus cudaMemcpyPeerAsync for comparison.  Needs >= 2 GPUs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bcast_bw tools/ubench/bcast_bw.cu && /tmp/bcast_bw
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>
struct P { const float4* src; float4* dst[8]; int ndst; long long n4; };
template <int U>
__global__ void __launch_bounds__(256) bcast(P p) {
    const long long stride = (long long)gridDim.x * blockDim.x * U;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x); i < p.n4; i += stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long j = i + (long long)u * gridDim.x * blockDim.x;
            if (j < p.n4) v[u] = __ldg(p.src + j);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long j = i + (long long)u * gridDim.x * blockDim.x;
            if (j < p.n4)
                for (int d = 0; d < p.ndst; ++d) p.dst[d][j] = v[u];
        }
    }
    __threadfence_system();
}
int main() {
    int n = 0; cudaGetDeviceCount(&n);
    if (n < 2) { printf("need 2 GPUs\n"); return 0; }
    const int npeer = n - 1 < 7 ? n - 1 : 7;
    const size_t bytes = 2560ull << 10;      // 2.5 MB per destination
    cudaSetDevice(0);
    float4* src; cudaMalloc(&src, bytes); cudaMemset(src, 1, bytes);
    std::vector<float4*> dst(7);
    for (int d = 0; d < 7; ++d) {
        const int dev = 1 + d % npeer;
        cudaSetDevice(dev); cudaMalloc(&dst[d], bytes);
        cudaSetDevice(0); cudaDeviceEnablePeerAccess(dev, 0);
    }
    cudaSetDevice(0); cudaGetLastError();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int ndst : {1, 7}) for (int grid : {32, 64, 128, 256}) for (int U : {1, 4}) {
        P p{}; p.src = src; p.ndst = ndst; p.n4 = bytes / 16; for (int d = 0; d < ndst; ++d) p.dst[d] = dst[d];
        for (int rep = 0; rep < 12; ++rep) {
            if (rep == 2) cudaEventRecord(a);
            if (U == 1) bcast<1><<<grid, 256>>>(p); else bcast<4><<<grid, 256>>>(p);
        }
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10;
        printf("kernel ndst %d (over %d peers) grid %3d unroll %d: %.1f us per launch, %.1f GB/s egress\n", ndst, npeer < ndst ? npeer : ndst, grid, U, ms * 1e3, ndst * bytes / (ms * 1e-3) / 1e9);
    }
    for (int ndst : {1, 7}) {
        for (int rep = 0; rep < 12; ++rep) {
            if (rep == 2) cudaEventRecord(a);
            for (int d = 0; d < ndst; ++d) cudaMemcpyPeerAsync(dst[d], 1 + d % npeer, src, 0, bytes, 0);
        }
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10;
        printf("cudaMemcpyPeerAsync x %d on one stream: %.1f us, %.1f GB/s egress\n", ndst, ms * 1e3, ndst * bytes / (ms * 1e-3) / 1e9);
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
