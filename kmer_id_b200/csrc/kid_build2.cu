// kid_build2.cu - GPU-side construction of the minimizer-addressed table (kid_table2.cuh).
//
// Reproduces Hashtable::add_kmer's observable semantics (newkmer_10nx.cpp:235-263: the first line
// of a key wins, taxon-0 lines are invisible) without depending on the packed sector format:
//   1. CLAIM   every probe finds-or-claims a slot for its key in an `owner` array (one uint32 per
//              slot = index of a probe line): atomicCAS on an empty slot, or, if the slot's owner
//              carries the same key, atomicMin so that the LOWEST file index ends up owning it;
//   2. PACK    every sector gathers its three owners and writes keys + taxa in the packed format.
// Slots of a sector fill left to right and sectors only ever fill up, which is the invariant the
// lookup relies on (stop at the first sector that has an empty slot).
#include "kid_kernels.cuh"

namespace {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;

__global__ void __launch_bounds__(256)
kid2_claim_kernel(uint32_t *owner, int line_shift, uint64_t sector_mask, const uint64_t *__restrict__ keys,
                  const uint32_t *__restrict__ taxa, size_t n_keys, uint32_t n_taxa, Kid2BuildStatus *status)
{
    unsigned long long claimed = 0, displaced = 0;
    unsigned max_probe = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_keys;
         i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t t = taxa[i];
        if (t == 0) continue; // value == 0 means "empty" in the reference: never visible
        if (t >= n_taxa || t > KID2_MAX_TAXA) { status->range_error = 1; continue; }
        const uint64_t key = keys[i] & KID_MASK60;
        const uint64_t home = kid2_home_sector(kid_minimizer(key), key, line_shift);
        bool placed = false;
        for (unsigned d = 0; d <= KID2_BUILD_MAX_PROBE && !placed; d++) {
            const uint64_t sec = (home + d) & sector_mask;
            for (int j = 0; j < KID2_SLOTS_PER_SECTOR; j++) {
                uint32_t *op = owner + KID2_SLOTS_PER_SECTOR * sec + j;
                uint32_t o = *reinterpret_cast<volatile uint32_t *>(op);
                if (o == kEmpty) {
                    o = atomicCAS(op, kEmpty, (uint32_t)i);
                    if (o == kEmpty) { // we opened the slot
                        claimed++;
                        displaced += d > 0;
                        max_probe = max(max_probe, d);
                        placed = true;
                        break;
                    }
                }
                if ((keys[o] & KID_MASK60) == key) { // same key: the earliest line keeps the slot
                    atomicMin(op, (uint32_t)i);
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
kid2_pack_kernel(uint4 *sectors, size_t n_sectors, const uint32_t *__restrict__ owner,
                 const uint64_t *__restrict__ keys, const uint32_t *__restrict__ taxa)
{
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_sectors;
         s += (size_t)gridDim.x * blockDim.x) {
        uint32_t w[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        uint64_t tx = 0;
#pragma unroll
        for (int j = 0; j < KID2_SLOTS_PER_SECTOR; j++) {
            const uint32_t o = owner[KID2_SLOTS_PER_SECTOR * s + j];
            if (o == kEmpty) continue;
            const uint64_t kw = (keys[o] & KID_MASK60) | KID2_OCC;
            w[2 * j] = (uint32_t)kw;
            w[2 * j + 1] = (uint32_t)(kw >> 32);
            tx |= (uint64_t)taxa[o] << (KID2_TAXON_BITS * j); // the first file line of this key
        }
        w[6] = (uint32_t)tx;
        w[7] = (uint32_t)(tx >> 32);
        sectors[2 * s] = make_uint4(w[0], w[1], w[2], w[3]);
        sectors[2 * s + 1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

} // namespace

cudaError_t kid_launch_build2(uint4 *sectors, int log2_lines, uint32_t *owner, const uint64_t *keys,
                              const uint32_t *taxa, size_t n_keys, int n_taxa, Kid2BuildStatus *status,
                              cudaStream_t stream)
{
    const uint64_t n_sectors = 4ULL << log2_lines;
    if (n_keys) {
        kid2_claim_kernel<<<148 * 16, 256, 0, stream>>>(owner, 32 - log2_lines, n_sectors - 1, keys, taxa, n_keys,
                                                        (uint32_t)n_taxa, status);
        KID_COUNT_LAUNCH();
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) return err;
    }
    kid2_pack_kernel<<<148 * 16, 256, 0, stream>>>(sectors, n_sectors, owner, keys, taxa);
    KID_COUNT_LAUNCH();
    return cudaGetLastError();
}
