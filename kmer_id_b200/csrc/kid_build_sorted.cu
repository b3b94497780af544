// kid_build_sorted.cu - deterministic GPU construction of the probe table (both layouts).
//
// Replaces Hashtable::add_kmer (newkmer_10nx.cpp:235-263).  Observable semantics reproduced: the
// FIRST file line of a key wins, lines with taxon 0 are invisible.
//
// The placement must not depend on thread timing: in a multi-GPU run every rank builds its own
// replica and the per-slot "seen" bitmaps are OR-ed across ranks BY SLOT INDEX, so the same key has
// to land in the same slot everywhere.  Hence sort + scan instead of racing atomicCAS claims:
//   1. stable radix sort of (key, line index)      -> equal keys adjacent, earliest line first
//   2. run heads = distinct keys, owner = that earliest line (first wins)
//   3. stable radix sort of heads by home sector   -> probe order inside a cluster is (home, key)
//   4. slot_i = max(S*home_i, slot_{i-1} + 1)      = i + prefix-max(S*home_j - j): one scan
//      (linear probing in sorted order: every slot between a key's home and its slot is occupied,
//       which is the invariant the lookups rely on; the table has slack sectors instead of wrap)
//   5. owner[slot_i] = line index; a pack kernel then writes the sectors in the layout's format.
#include "kid_kernels.cuh"

#include <cub/cub.cuh>

namespace {

constexpr uint64_t kSentinel = 1ULL << 60; // sorts after every real 60-bit key
constexpr uint32_t kEmpty = 0xFFFFFFFFu;

__global__ void __launch_bounds__(256)
prep_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ taxa, size_t n, uint32_t n_taxa,
            uint32_t max_taxon, uint64_t *skey, uint32_t *sidx, unsigned int *range_error)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t t = taxa[i];
        if (t >= n_taxa || t > max_taxon) *range_error = 1;
        skey[i] = (t == 0 || t >= n_taxa || t > max_taxon) ? kSentinel : (keys[i] & KID_MASK60);
        sidx[i] = (uint32_t)i;
    }
}

// heads of key runs get their home sector, everything else the "no home" sentinel
__global__ void __launch_bounds__(256)
home_kernel(const uint64_t *__restrict__ skey, size_t n, KidSortedBuildParams p, uint64_t *home,
            unsigned long long *n_unique)
{
    unsigned long long cnt = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t k = skey[i];
        const bool head = k != kSentinel && (i == 0 || skey[i - 1] != k);
        uint64_t h = ~0ULL;
        if (head) {
            h = p.layout == KID_LAYOUT_KEYHASH ? (kid_hash60(k) >> p.rem_bits)
                                               : kid2_home_sector(kid_minimizer_mm(k, p.mm), k, p.line_shift, p.sub_bits);
            cnt++;
        }
        home[i] = h;
    }
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(n_unique, cnt);
}

__global__ void __launch_bounds__(256)
gap_kernel(const uint64_t *__restrict__ home, size_t n_u, int S, long long *q)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_u; i += (size_t)gridDim.x * blockDim.x)
        q[i] = (long long)(home[i] * (uint64_t)S) - (long long)i;
}

__global__ void __launch_bounds__(256)
place_kernel(const uint64_t *__restrict__ home, const long long *__restrict__ m, const uint32_t *__restrict__ line,
             size_t n_u, int S, uint64_t total_slots, uint32_t max_disp, uint32_t *owner, Kid2BuildStatus *st)
{
    unsigned long long displaced = 0;
    unsigned maxd = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_u; i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t slot = (uint64_t)(m[i] + (long long)i);
        const uint64_t d = slot / (uint64_t)S - home[i];
        if (slot >= total_slots || d > max_disp) { st->overflow = 1; continue; }
        owner[slot] = line[i];
        displaced += d > 0;
        maxd = max(maxd, (unsigned)d);
    }
    for (int o = 16; o; o >>= 1) {
        displaced += __shfl_xor_sync(0xFFFFFFFFu, displaced, o);
        maxd = max(maxd, __shfl_xor_sync(0xFFFFFFFFu, maxd, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (displaced) atomicAdd(&st->n_displaced, displaced);
        if (maxd) atomicMax(&st->max_probe, maxd);
    }
}

struct MaxOp {
    __device__ __forceinline__ long long operator()(long long a, long long b) const { return a > b ? a : b; }
};

// ---- pack kernels: owner[] -> sectors in the layout's format ----------------------------------
__global__ void __launch_bounds__(256)
pack2_kernel(uint4 *sectors, size_t n_sectors, const uint32_t *__restrict__ owner,
             const uint64_t *__restrict__ keys, const uint32_t *__restrict__ taxa)
{
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_sectors; s += (size_t)gridDim.x * blockDim.x) {
        uint32_t w[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        uint64_t tx = 0;
#pragma unroll
        for (int j = 0; j < KID2_SLOTS_PER_SECTOR; j++) {
            const uint32_t o = owner[KID2_SLOTS_PER_SECTOR * s + j];
            if (o == kEmpty) continue;
            const uint64_t kw = (keys[o] & KID_MASK60) | KID2_OCC;
            w[j] = (uint32_t)(kw >> 32); // first half: high words (kid_table2.cuh)
            w[4 + j] = (uint32_t)kw;     // second half: low words
            tx |= (uint64_t)taxa[o] << (KID2_TAXON_BITS * j); // taxon of the first file line of this key
        }
        w[3] = (uint32_t)tx;
        w[7] = (uint32_t)(tx >> 32);
        sectors[2 * s] = make_uint4(w[0], w[1], w[2], w[3]);
        sectors[2 * s + 1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

__global__ void __launch_bounds__(256)
pack1_kernel(uint64_t *slots, size_t n_slots, int rem_bits, const uint32_t *__restrict__ owner,
             const uint64_t *__restrict__ keys, const uint32_t *__restrict__ taxa)
{
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += (size_t)gridDim.x * blockDim.x) {
        const uint32_t o = owner[s];
        uint64_t e = 0;
        if (o != kEmpty) {
            const uint64_t h = kid_hash60(keys[o] & KID_MASK60);
            const uint64_t disp = (uint64_t)(s >> 2) - (h >> rem_bits);
            const uint64_t rem = h & ((1ULL << rem_bits) - 1ULL);
            e = (((rem << KID_DISP_BITS) | disp) << KID_TAXON_BITS) | (uint64_t)taxa[o];
        }
        slots[s] = e;
    }
}

} // namespace

#define KID_TRY(call)                         \
    do {                                      \
        cudaError_t e_ = (call);              \
        if (e_ != cudaSuccess) { cleanup(); return e_; } \
    } while (0)

cudaError_t kid_build_owner_sorted(const uint64_t *keys, const uint32_t *taxa, size_t n, const KidSortedBuildParams &p,
                                   uint32_t *owner, Kid2BuildStatus *dstatus, Kid2BuildStatus *hstatus,
                                   cudaStream_t stream)
{
    const uint64_t total_slots = (p.n_sectors + p.slack_sectors) * (uint64_t)p.slots_per_sector;
    uint64_t *a64 = nullptr, *b64 = nullptr;
    uint32_t *a32 = nullptr, *b32 = nullptr;
    void *tmp = nullptr;
    unsigned long long *d_nu = nullptr;
    auto cleanup = [&]() { cudaFree(a64); cudaFree(b64); cudaFree(a32); cudaFree(b32); cudaFree(tmp); cudaFree(d_nu); };
    KID_TRY(cudaMemsetAsync(owner, 0xFF, total_slots * sizeof(uint32_t), stream));
    KID_TRY(cudaMemsetAsync(dstatus, 0, sizeof(Kid2BuildStatus), stream));
    memset(hstatus, 0, sizeof *hstatus);
    if (n == 0) return cudaStreamSynchronize(stream);

    KID_TRY(cudaMalloc(&a64, n * 8));
    KID_TRY(cudaMalloc(&b64, n * 8));
    KID_TRY(cudaMalloc(&a32, n * 4));
    KID_TRY(cudaMalloc(&b32, n * 4));
    KID_TRY(cudaMalloc(&d_nu, 8));
    KID_TRY(cudaMemsetAsync(d_nu, 0, 8, stream));
    size_t tb1 = 0, tb2 = 0, tb3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb1, a64, b64, a32, b32, n, 0, 61, stream);
    cub::DeviceRadixSort::SortPairs(nullptr, tb2, b64, a64, b32, a32, n, 0, 64, stream);
    cub::DeviceScan::InclusiveScan(nullptr, tb3, (long long *)b64, (long long *)b64, MaxOp(), n, stream);
    const size_t tb = tb1 > tb2 ? (tb1 > tb3 ? tb1 : tb3) : (tb2 > tb3 ? tb2 : tb3);
    KID_TRY(cudaMalloc(&tmp, tb ? tb : 16));
    const unsigned grid = 148 * 16;

    // 1. (key, line) sorted by key, stable -> earliest line first inside a run
    prep_kernel<<<grid, 256, 0, stream>>>(keys, taxa, n, p.n_taxa, p.max_taxon, a64, a32, &dstatus->range_error);
    KID_COUNT_LAUNCH();
    size_t t = tb;
    KID_TRY(cub::DeviceRadixSort::SortPairs(tmp, t, a64, b64, a32, b32, n, 0, 61, stream));
    // 2. heads -> home sector (others: sentinel), count distinct keys
    home_kernel<<<grid, 256, 0, stream>>>(b64, n, p, a64, d_nu);
    KID_COUNT_LAUNCH();
    // 3. stable sort by home sector: heads first, clusters in (home, key) order
    t = tb;
    KID_TRY(cub::DeviceRadixSort::SortPairs(tmp, t, a64, b64, b32, a32, n, 0, 64, stream));
    unsigned long long n_u = 0;
    KID_TRY(cudaMemcpyAsync(&n_u, d_nu, 8, cudaMemcpyDeviceToHost, stream));
    KID_TRY(cudaStreamSynchronize(stream));
    // now b64[0..n_u) = home sectors ascending, a32[0..n_u) = owning line of each distinct key
    if (n_u) {
        // 4. slot_i = i + prefix-max(S*home_j - j)
        gap_kernel<<<grid, 256, 0, stream>>>(b64, (size_t)n_u, p.slots_per_sector, (long long *)a64);
        KID_COUNT_LAUNCH();
        t = tb;
        KID_TRY(cub::DeviceScan::InclusiveScan(tmp, t, (long long *)a64, (long long *)a64, MaxOp(), (size_t)n_u, stream));
        // 5. scatter the owners
        place_kernel<<<grid, 256, 0, stream>>>(b64, (const long long *)a64, a32, (size_t)n_u, p.slots_per_sector,
                                               total_slots, p.max_disp, owner, dstatus);
        KID_COUNT_LAUNCH();
    }
    KID_TRY(cudaMemcpyAsync(hstatus, dstatus, sizeof *hstatus, cudaMemcpyDeviceToHost, stream));
    KID_TRY(cudaStreamSynchronize(stream));
    hstatus->n_distinct = n_u;
    cleanup();
    return cudaGetLastError();
}

cudaError_t kid_launch_pack2(uint4 *sectors, size_t n_sectors_total, const uint32_t *owner, const uint64_t *keys,
                             const uint32_t *taxa, cudaStream_t stream)
{
    pack2_kernel<<<148 * 16, 256, 0, stream>>>(sectors, n_sectors_total, owner, keys, taxa);
    KID_COUNT_LAUNCH();
    return cudaGetLastError();
}

cudaError_t kid_launch_pack1(uint64_t *slots, size_t n_slots_total, int rem_bits, const uint32_t *owner,
                             const uint64_t *keys, const uint32_t *taxa, cudaStream_t stream)
{
    pack1_kernel<<<148 * 16, 256, 0, stream>>>(slots, n_slots_total, rem_bits, owner, keys, taxa);
    KID_COUNT_LAUNCH();
    return cudaGetLastError();
}
