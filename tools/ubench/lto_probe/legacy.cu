#include <cuda_runtime.h>
#include <cufft.h>
#include <cufftXt.h>
#include <stdio.h>
__device__ cufftComplex ld_cb(void* dataIn, size_t offset, void* callerInfo, void* sharedPtr) { return ((cufftComplex*)callerInfo)[offset & 1023]; }
__device__ void st_cb(void* dataOut, size_t offset, cufftComplex e, void* callerInfo, void* sharedPtr) { atomicAdd((float*)callerInfo + (offset & 1023), e.x * e.x + e.y * e.y); }
__device__ cufftCallbackLoadC d_ld = ld_cb;
__device__ cufftCallbackStoreC d_st = st_cb;
extern "C" int probe(int N, int batch) {
    cufftHandle p; cufftCreate(&p); size_t ws;
    cufftResult r = cufftMakePlan1d(p, N, CUFFT_C2C, batch, &ws); printf("makeplan %d\n", r);
    cufftCallbackLoadC h_ld; cufftCallbackStoreC h_st;
    cudaMemcpyFromSymbol(&h_ld, d_ld, sizeof(h_ld)); cudaMemcpyFromSymbol(&h_st, d_st, sizeof(h_st));
    void* info; cudaMalloc(&info, 1 << 16);
    r = cufftXtSetCallback(p, (void**)&h_ld, CUFFT_CB_LD_COMPLEX, &info); printf("set ld %d\n", r);
    r = cufftXtSetCallback(p, (void**)&h_st, CUFFT_CB_ST_COMPLEX, &info); printf("set st %d\n", r);
    return 0;
}
int main() { return probe(4096, 8); }
