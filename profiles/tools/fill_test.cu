// write-only bandwidth probe: how fast can 1.31 GB be filled with 16-byte stores?
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void fill(double2* p, size_t n16, int per) {
    size_t i = (size_t)blockIdx.x * blockDim.x * per + threadIdx.x;
    double2 v = make_double2(1.0, 1.0);
#pragma unroll 8
    for (int j = 0; j < per; ++j, i += blockDim.x) {
        if (i < n16) {
            if (MODE == 0) p[i] = v;
            else if (MODE == 1) __stcs(p + i, v);
            else if (MODE == 2) __stwt(p + i, v);
            else __stcg(p + i, v);
        }
    }
}
int main() {
    size_t bytes = 8192ull * 20000 * 8, n16 = bytes / 16;
    double2* p; cudaMalloc(&p, bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int mode = 0; mode < 4; ++mode)
        for (int per : {1, 8, 32}) {
            int threads = 256; size_t blocks = (n16 + (size_t)threads * per - 1) / ((size_t)threads * per);
            auto run = [&]() {
                if (mode == 0) fill<0><<<blocks, threads>>>(p, n16, per);
                if (mode == 1) fill<1><<<blocks, threads>>>(p, n16, per);
                if (mode == 2) fill<2><<<blocks, threads>>>(p, n16, per);
                if (mode == 3) fill<3><<<blocks, threads>>>(p, n16, per);
            };
            for (int w = 0; w < 3; ++w) run();
            cudaEventRecord(a);
            for (int r = 0; r < 20; ++r) run();
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); ms /= 20;
            printf("mode %d per %2d: %.3f ms  %.0f GB/s\n", mode, per, ms, bytes / ms / 1e6);
        }
    cudaEventRecord(a);
    for (int r = 0; r < 20; ++r) cudaMemsetAsync(p, 0, bytes);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 20;
    printf("cudaMemset: %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
    return 0;
}
