// kid_build2.cu - GPU-side construction of the minimizer-addressed table (kid_table2.cuh).
//
// Same two-phase scheme as kid_build.cu, reproducing Hashtable::add_kmer's observable semantics
// (newkmer_10nx.cpp:235-263: first line of a key wins, taxon-0 lines are invisible):
//   1. every probe finds-or-claims the slot of its key by atomicCAS on the 64-bit key word, walking
//      sectors from its home, and keeps the lowest file index in the slot's aux word (atomicMin);
//   2. every claimed slot takes the taxon of that lowest index.
#include "kid_kernels.cuh"

namespace {

__global__ void __launch_bounds__(256) kid2_fill_kernel(uint4 *entries, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        entries[i] = make_uint4(0u, 0u, 0u, 0xFFFFFFFFu);
}

__global__ void __launch_bounds__(256)
kid2_claim_kernel(uint4 *entries, int line_shift, uint64_t sector_mask, const uint64_t *__restrict__ keys,
                  const uint32_t *__restrict__ taxa, size_t n_keys, uint32_t n_taxa, Kid2BuildStatus *status)
{
    unsigned long long claimed = 0, displaced = 0;
    unsigned max_probe = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_keys;
         i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t t = taxa[i];
        if (t == 0) continue; // value == 0 means "empty" in the reference: never visible
        if (t >= n_taxa) { status->range_error = 1; continue; }
        const uint64_t key = keys[i] & KID_MASK60;
        const unsigned long long want = key | KID2_OCC;
        const uint64_t home = kid2_home_sector(kid_minimizer(key), key, line_shift);
        bool placed = false;
        for (unsigned d = 0; d <= KID2_BUILD_MAX_PROBE && !placed; d++) {
            const uint64_t sec = (home + d) & sector_mask;
            for (int j = 0; j < 2; j++) {
                uint4 *ep = entries + 2 * sec + j;
                unsigned long long *kp = reinterpret_cast<unsigned long long *>(ep);
                unsigned long long e = *reinterpret_cast<volatile unsigned long long *>(kp);
                if (e == 0) {
                    e = atomicCAS(kp, 0ULL, want);
                    if (e == 0) { claimed++; displaced += d > 0; e = want; }
                }
                if (e == want) {
                    atomicMin(reinterpret_cast<unsigned int *>(ep) + 3, (unsigned int)i);
                    max_probe = max(max_probe, d);
                    placed = true;
                    break;
                }
            }
        }
        if (!placed) status->overflow = 1;
    }
    for (int o = 16; o; o >>= 1) {
        claimed += __shfl_xor_sync(0xFFFFFFFFu, claimed, o);
        displaced += __shfl_xor_sync(0xFFFFFFFFu, displaced, o);
        max_probe = max(max_probe, __shfl_xor_sync(0xFFFFFFFFu, max_probe, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (claimed) atomicAdd(&status->n_distinct, claimed);
        if (displaced) atomicAdd(&status->n_displaced, displaced);
        if (max_probe) atomicMax(&status->max_probe, max_probe);
    }
}

__global__ void __launch_bounds__(256)
kid2_resolve_kernel(uint4 *entries, size_t n_entries, const uint32_t *__restrict__ taxa)
{
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_entries;
         s += (size_t)gridDim.x * blockDim.x) {
        uint4 e = entries[s];
        if ((e.x | e.y) == 0) continue;
        e.z = taxa[e.w]; // the first file line that carried this key
        e.w = 0;
        entries[s] = e;
    }
}

} // namespace

cudaError_t kid_launch_fill2(uint4 *entries, size_t n_entries, cudaStream_t stream)
{
    if (n_entries == 0) return cudaSuccess;
    kid2_fill_kernel<<<148 * 16, 256, 0, stream>>>(entries, n_entries);
    KID_COUNT_LAUNCH();
    return cudaGetLastError();
}

cudaError_t kid_launch_build2(uint4 *entries, int log2_lines, const uint64_t *keys, const uint32_t *taxa,
                              size_t n_keys, int n_taxa, Kid2BuildStatus *status, cudaStream_t stream)
{
    if (n_keys == 0) return cudaSuccess;
    const uint64_t n_sectors = 4ULL << log2_lines;
    kid2_claim_kernel<<<148 * 16, 256, 0, stream>>>(entries, 32 - log2_lines, n_sectors - 1, keys, taxa, n_keys,
                                                    (uint32_t)n_taxa, status);
    KID_COUNT_LAUNCH();
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    kid2_resolve_kernel<<<148 * 16, 256, 0, stream>>>(entries, 2 * n_sectors, taxa);
    KID_COUNT_LAUNCH();
    return cudaGetLastError();
}
