// variants of fill_test2: who writes a chunk -- one warp (k_rr_points today), or the 8 warps of a CTA together
// (warp w takes the rows w, w+8, ... of 512 bytes).   nvcc -arch=sm_100a -O3 -o fill_test3 fill_test3.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int COOP>
__global__ void __launch_bounds__(256, 3) fill_items(double2 *p, size_t n16, int chunk16, int *work) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long nitems = (long long)((n16 + chunk16 - 1) / chunk16);
    const double2 v = make_double2(1.0, 1.0);
    __shared__ long long s_it;
    for (;;) {
        long long it = 0;
        if (COOP) {
            __syncthreads();
            if (threadIdx.x == 0) s_it = atomicAdd(work, 1);
            __syncthreads();
            it = s_it;
        } else {
            if (lane == 0) it = atomicAdd(work, 1);
            it = __shfl_sync(0xffffffffu, it, 0);
        }
        if (it >= nitems) break;
        const size_t base = (size_t)it * chunk16;
        const size_t end = ((size_t)(it + 1) * chunk16 < n16 ? (size_t)(it + 1) * chunk16 : n16) - base;
        if (COOP) {   // rows of 32 double2 (512 B): warp w writes rows w, w+8, ...; 4 rows in flight per warp
            for (size_t r = warp; r * 32 < end; r += 32) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const size_t o = (r + j * 8) * 32 + lane;
                    if (o < end) __stcs(p + base + o, v);
                }
            }
        } else {
            double2 *q = p + base + lane;
            for (size_t o = 0; o < end; o += 256) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (o + j * 32 + lane < end) __stcs(q + o + j * 32, v);
            }
        }
    }
}
int main() {
    size_t bytes = 8192ull * 20000 * 8, n16 = bytes / 16;
    double2 *p; cudaMalloc(&p, bytes);
    int *work; cudaMalloc(&work, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int coop = 0; coop < 2; ++coop)
        for (int chunk_kb : {40, 160, 320, 640}) {
            const int chunk16 = chunk_kb * 1024 / 16;
            float best = 1e9;
            for (int r = 0; r < 12; ++r) {
                cudaMemsetAsync(work, 0, 4);
                cudaEventRecord(a);
                if (coop) fill_items<1><<<148 * 3, 256>>>(p, n16, chunk16, work);
                else fill_items<0><<<148 * 3, 256>>>(p, n16, chunk16, work);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b);
                if (r >= 2 && ms < best) best = ms;
            }
            printf("%s, %3d KB chunks: %.3f ms  %.0f GB/s\n", coop ? "CTA-cooperative" : "warp-private   ", chunk_kb, best, bytes / best / 1e6);
        }
    return 0;
}
