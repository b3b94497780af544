// gather_bench.cu - practical ceiling for the probe-table access pattern: independent random
// 32-byte sector reads over a table of a given size (no reuse, no input stream).
//   ./gather_bench <table_MiB> [unroll=4] [threads_per_block=256] [blocks_per_sm=8] [sectors_per_access=1] [l2_fetch_granularity=0]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

template <int U, int W>
__global__ void gather(const uint64_t *__restrict__ tab, uint64_t mask, int iters, uint64_t *out)
{
    uint64_t acc = 0;
    uint64_t ctr = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x100000001ULL;
    for (int it = 0; it < iters; it++) {
        uint64_t e[U][W][4];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t b = (mix(ctr++) & mask) & ~(uint64_t)(W - 1); // W sectors, W*32-B aligned
#pragma unroll
            for (int w = 0; w < W; w++)
                asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                             : "=l"(e[u][w][0]), "=l"(e[u][w][1]), "=l"(e[u][w][2]), "=l"(e[u][w][3])
                             : "l"(tab + 4 * (b + w)));
        }
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int w = 0; w < W; w++) acc ^= e[u][w][0] ^ e[u][w][1] ^ e[u][w][2] ^ e[u][w][3];
    }
    if (acc == 0x1234567) out[0] = acc;
}

int main(int argc, char **argv)
{
    const size_t mib = argc > 1 ? strtoull(argv[1], 0, 10) : 4096;
    const int unroll = argc > 2 ? atoi(argv[2]) : 4;
    const int tpb = argc > 3 ? atoi(argv[3]) : 256;
    const int bps = argc > 4 ? atoi(argv[4]) : 8;
    const int width = argc > 5 ? atoi(argv[5]) : 1;   // sectors per access: 1, 2 or 4
    const int gran = argc > 6 ? atoi(argv[6]) : 0;    // cudaLimitMaxL2FetchGranularity (0 = leave)
    if (gran) {
        cudaError_t ge = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("L2 fetch granularity: asked %d -> %zu (%s)\n", gran, got, cudaGetErrorString(ge));
    } else {
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("L2 fetch granularity default: %zu\n", got);
    }
    const size_t bytes = mib << 20;
    uint64_t n_buckets = bytes / 32;
    uint64_t mask = 1;
    while (mask * 2 <= n_buckets) mask *= 2;
    mask -= 1;
    uint64_t *tab, *out;
    if (cudaMalloc(&tab, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&out, 8);
    cudaMemset(tab, 1, bytes);
    const int iters = 64;
    const int nsm = argc > 7 ? atoi(argv[7]) : 148;
    const int grid = nsm * bps;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(a);
        if (width == 1) {
            if (unroll == 1) gather<1, 1><<<grid, tpb>>>(tab, mask, iters, out);
            else if (unroll == 2) gather<2, 1><<<grid, tpb>>>(tab, mask, iters, out);
            else if (unroll == 4) gather<4, 1><<<grid, tpb>>>(tab, mask, iters, out);
            else gather<8, 1><<<grid, tpb>>>(tab, mask, iters, out);
        } else if (width == 2) {
            if (unroll <= 2) gather<2, 2><<<grid, tpb>>>(tab, mask, iters, out);
            else gather<4, 2><<<grid, tpb>>>(tab, mask, iters, out);
        } else {
            if (unroll <= 2) gather<2, 4><<<grid, tpb>>>(tab, mask, iters, out);
            else gather<4, 4><<<grid, tpb>>>(tab, mask, iters, out);
        }
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    const double n = (double)grid * tpb * iters * unroll;
    printf("table %6zu MiB unroll %2d tpb %4d bps %2d width %d : %8.3f ms  %7.2f G accesses/s  %7.1f GB/s useful  (%s)\n",
           mib, unroll, tpb, bps, width, best, n / best / 1e6, n * 32 * width / best / 1e6, cudaGetErrorString(e));
    return 0;
}
