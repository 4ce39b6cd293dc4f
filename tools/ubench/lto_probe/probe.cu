#include <cuda_runtime.h>
#include <cufft.h>
#include <cufftXt.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
int main(int argc, char** argv) {
    const char* path = argv[1]; const char* ld = argv[2]; const char* st = argv[3];
    int N = atoi(argv[4]), batch = atoi(argv[5]); int mode = atoi(argv[6]);   // 1 load, 2 store, 3 both
    FILE* f = fopen(path, "rb"); fseek(f, 0, SEEK_END); size_t sz = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<char> buf(sz); fread(buf.data(), 1, sz, f); fclose(f);
    cudaFree(0);
    void* d_info; cudaMalloc(&d_info, 256); cudaMemset(d_info, 0, 256);
    cufftHandle p; cufftResult r = cufftCreate(&p); printf("create %d\n", r);
    if (mode & 1) { r = cufftXtSetJITCallback(p, ld, buf.data(), sz, CUFFT_CB_LD_COMPLEX, &d_info); printf("set load %s -> %d\n", ld, r); }
    if (mode & 2) { r = cufftXtSetJITCallback(p, st, buf.data(), sz, CUFFT_CB_ST_COMPLEX, &d_info); printf("set store %s -> %d\n", st, r); }
    size_t ws = 0; r = cufftMakePlan1d(p, N, CUFFT_C2C, batch, &ws); printf("makeplan N=%d batch=%d mode=%d -> %d (ws %zu)\n", N, batch, mode, r, ws);
    int v; cufftGetVersion(&v); printf("cufft version %d\n", v);
    return 0;
}
