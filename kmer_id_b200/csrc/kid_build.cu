// kid_build.cu - GPU-side construction of the probe table.
//
// Replaces Hashtable::add_kmer (newkmer_10nx.cpp:235-263).  The reference inserts serially in
// file order and never checks for an equal key, so a later duplicate is shadowed by the earlier
// one ("first wins"), and a probe whose target is 0 leaves its cell looking empty, i.e. it is
// invisible.  A parallel build cannot rely on arrival order, so it runs in two
// phases:
//   1. every probe i with taxa[i] != 0 finds-or-claims the slot of its key (atomicCAS on an empty
//      slot, placeholder taxon) and records owner[slot] = min(owner[slot], i);
//   2. every claimed slot takes the taxon of its owner - the lowest file index, i.e. the line the
//      reference would have found first.
// Slots only ever go from empty to occupied, which keeps the invariant the lookup relies on:
// an entry at displacement d implies buckets home .. home+d-1 were already full.
#include "kid_kernels.cuh"

namespace {

__global__ void __launch_bounds__(256)
kid_build_claim_kernel(uint64_t *slots, int rem_bits, uint64_t bucket_mask, uint32_t *owner,
                       const uint64_t *__restrict__ keys, const uint32_t *__restrict__ taxa,
                       size_t n_keys, uint32_t n_taxa, KidBuildStatus *status)
{
    unsigned long long claimed = 0, displaced = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_keys;
         i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t t = taxa[i];
        if (t == 0) continue; // value == 0 means "empty" in the reference: never visible
        if (t >= n_taxa || t >= KID_TAXON_MASK) { status->range_error = 1; continue; }
        const uint64_t h = kid_hash60(keys[i] & KID_MASK60);
        const uint64_t home = h >> rem_bits;
        const uint64_t rem = h & ((1ULL << rem_bits) - 1ULL);
        bool placed = false;
        for (int d = 0; d <= KID_MAX_DISP && !placed; d++) {
            const uint64_t base = 4 * ((home + (uint64_t)d) & bucket_mask);
            const uint64_t tag = (rem << KID_DISP_BITS) | (uint64_t)d;
            const uint64_t fresh = (tag << KID_TAXON_BITS) | KID_TAXON_MASK; // placeholder taxon
            for (int j = 0; j < 4; j++) {
                unsigned long long *sp = reinterpret_cast<unsigned long long *>(slots + base + j);
                unsigned long long e = *reinterpret_cast<volatile unsigned long long *>(sp);
                if (e == 0) {
                    e = atomicCAS(sp, 0ULL, (unsigned long long)fresh);
                    if (e == 0) { // we created the entry
                        claimed++;
                        displaced += d > 0;
                        e = fresh;
                    }
                }
                if ((e >> KID_TAXON_BITS) == tag) {
                    atomicMin(owner + base + j, (uint32_t)i);
                    placed = true;
                    break;
                }
            }
        }
        if (!placed) status->overflow = 1;
    }
    // warp-aggregate the two counters
    for (int o = 16; o; o >>= 1) {
        claimed += __shfl_xor_sync(0xFFFFFFFFu, claimed, o);
        displaced += __shfl_xor_sync(0xFFFFFFFFu, displaced, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (claimed) atomicAdd(&status->n_distinct, claimed);
        if (displaced) atomicAdd(&status->n_displaced, displaced);
    }
}

__global__ void __launch_bounds__(256)
kid_build_resolve_kernel(uint64_t *slots, size_t n_slots, const uint32_t *__restrict__ owner,
                         const uint32_t *__restrict__ taxa)
{
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_slots;
         s += (size_t)gridDim.x * blockDim.x) {
        const uint64_t e = slots[s];
        if (e == 0) continue;
        slots[s] = (e & ~(uint64_t)KID_TAXON_MASK) | (uint64_t)taxa[owner[s]];
    }
}

} // namespace

cudaError_t kid_launch_build(uint64_t *slots, int log2_buckets, uint32_t *owner,
                             const uint64_t *keys, const uint32_t *taxa, size_t n_keys, int n_taxa,
                             KidBuildStatus *status, cudaStream_t stream)
{
    if (n_keys == 0) return cudaSuccess;
    const int rem_bits = 60 - log2_buckets;
    const uint64_t n_buckets = 1ULL << log2_buckets;
    const unsigned grid = 148 * 16;
    kid_build_claim_kernel<<<grid, 256, 0, stream>>>(slots, rem_bits, n_buckets - 1, owner, keys,
                                                      taxa, n_keys, (uint32_t)n_taxa, status);
    KID_COUNT_LAUNCH();
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    kid_build_resolve_kernel<<<grid, 256, 0, stream>>>(slots, 4 * n_buckets, owner, taxa);
    KID_COUNT_LAUNCH();
    return cudaGetLastError();
}
