// kid_classify.cu - the per-read hot path as one fused sm_100a kernel.
//
// Replaces, for a whole batch of reads, process_qual (newkmer_10nx.cpp:714-760), process_read
// (:452-617), Hashtable::getHash (:204-233) and Tree1::msca (:118-144).
//
// One warp owns one read at a time (persistent grid, warp-strided reads):
//   1. TRIM    quality trim with ballots: each of the reference's four `while` loops is a
//              "first/last position satisfying a predicate" search, done 32 positions per step.
//   2. STAGE   coalesced 128-bit loads of the read (aligned down to 16 B), each lane turns its 16
//              bases into one 32-bit word of 2-bit codes + a 16-bit ACGT-validity mask; both go
//              to a per-warp shared-memory strip (512 bases per window).
//   3. LOOKUP  lane j of chunk c owns k-mer start c*32+j: three LDS + funnel shifts give the
//              forward 60-bit key, __brevll gives the reverse complement, min() the canonical
//              key (:528).  Four chunks (128 k-mers) are hashed and their 32-byte buckets
//              requested before the first is consumed, so every lane keeps 4 independent DRAM
//              sectors in flight.
//   4. FOLD    hits are rare; a ballot finds them and the warp folds them strictly in position
//              order with kid_msca (the fold is order dependent).
//   5. COUNT   seen bit (atomicOr on the per-sample bitmap) for hits with taxon > 1 (:596-603),
//              gcount[final]++ (:613) in a shared-memory histogram flushed once per block.
#include "kid_kernels.cuh"

namespace {

constexpr int kWarpsPerBlock = KID_CLASSIFY1_THREADS / 32;
constexpr int kWindowStarts = 448; // k-mer starts served per window: 15 + 447 + 29 < 512
constexpr int kCodeWords = 36;     // 32 + slack for the 3-word read at the window end
constexpr int kValidWords = 20;    // 16 + slack
constexpr int kUnroll = 4;         // chunks of 32 k-mers in flight per warp

struct WarpStrip {
    uint32_t codes[kCodeWords];
    uint32_t valid[kValidWords];
};

// 4 ASCII bases in one 32-bit word (first base in the low byte) -> 8 bits of 2-bit codes with the
// first base in the top pair, and a 4-bit validity mask with the first base in the top bit.
__device__ __forceinline__ void pack4(uint32_t x, bool accept_u, uint32_t &code8, uint32_t &valid4)
{
    const uint32_t up = x & 0xDFDFDFDFu; // fold lower case onto upper case
    uint32_t ok = __vcmpeq4(up, 0x41414141u) | __vcmpeq4(up, 0x43434343u) |
                  __vcmpeq4(up, 0x47474747u) | __vcmpeq4(up, 0x54545454u);
    if (accept_u) ok |= __vcmpeq4(up, 0x55555555u);
    // (byte >> 1) & 3 : A->0 C->1 T,U->2 G->3 ; x ^ (x >> 1) swaps 2 and 3 -> A0 C1 G2 T3 (:480-519)
    uint32_t c = (x >> 1) & 0x03030303u;
    c ^= (c >> 1) & 0x01010101u;
    code8 = (c * 0x40100401u) >> 24;
    valid4 = ((ok & 0x01010101u) * 0x08040201u) >> 24 & 0xFu;
}

template <bool HAS_QUAL, bool SMEM_HIST>
__global__ void __launch_bounds__(KID_CLASSIFY1_THREADS, 3)
kid_classify_kernel(const KidClassifyParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpStrip *strips = reinterpret_cast<WarpStrip *>(smem_raw);
    int *hist = reinterpret_cast<int *>(smem_raw + sizeof(WarpStrip) * kWarpsPerBlock);

    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    WarpStrip &strip = strips[warp_in_block];
    const unsigned full = 0xFFFFFFFFu;

    if (SMEM_HIST) {
        for (int i = threadIdx.x; i < p.tree.n_taxa; i += blockDim.x) hist[i] = 0;
        __syncthreads();
    }
    if (lane < kCodeWords - 32) strip.codes[32 + lane] = 0;
    if (lane < kValidWords - 16) strip.valid[16 + lane] = 0;

    unsigned long long n_lookups = 0, n_hits = 0; // warp-uniform running totals

    const size_t warps_total = (size_t)gridDim.x * kWarpsPerBlock;
    for (size_t r = (size_t)blockIdx.x * kWarpsPerBlock + warp_in_block; r < p.n_reads;
         r += warps_total) {
        const uint64_t g0 = __ldg(p.off + r) - p.off_bias;
        const int len = (int)(__ldg(p.off + r + 1) - p.off_bias - g0);
        int start = 0, stop = len - 1;

        // ---------------------------------------------------------------- 1. TRIM (:724-753)
        if (HAS_QUAL && len > 0) {
            const signed char *q = reinterpret_cast<const signed char *>(p.qual) + g0;
            { // while (qual[start] < '1' && start < stop) start++;
                int ns = stop;
                for (int base = 0; base < stop; base += 32) {
                    const int i = base + lane;
                    const unsigned m = __ballot_sync(full, i < stop && q[i] >= 49);
                    if (m) { ns = base + __ffs(m) - 1; break; }
                }
                start = ns;
            }
            { // while (qual[stop] < '1' && stop > start) stop--;
                int nt = start;
                for (int hi = stop; hi > start; hi -= 32) {
                    const int i = hi - lane;
                    const unsigned m = __ballot_sync(full, i > start && q[i] >= 49);
                    if (m) { nt = hi - (__ffs(m) - 1); break; }
                }
                stop = nt;
            }
            if (start < stop - 4) { // leading 4-base window, sum(q-32) < 68 slides right
                const int lim = stop - 4;
                int ns = lim;
                for (int base = start; base < lim; base += 32) {
                    const int s = base + lane;
                    bool ok = false;
                    if (s < lim) ok = (int)q[s] + q[s + 1] + q[s + 2] + q[s + 3] - 128 >= 68;
                    const unsigned m = __ballot_sync(full, ok);
                    if (m) { ns = base + __ffs(m) - 1; break; }
                }
                start = ns;
            }
            if (start < stop - 4) { // trailing window slides left
                const int lo = start + 4;
                int nt = lo;
                for (int hi = stop; hi > lo; hi -= 32) {
                    const int t = hi - lane;
                    bool ok = false;
                    if (t > lo) ok = (int)q[t] + q[t - 1] + q[t - 2] + q[t - 3] - 128 >= 68;
                    const unsigned m = __ballot_sync(full, ok);
                    if (m) { nt = hi - (__ffs(m) - 1); break; }
                }
                stop = nt;
            }
        }
        if (p.out_span && lane == 0) {
            p.out_span[2 * r] = (uint32_t)start;
            p.out_span[2 * r + 1] = (uint32_t)stop;
        }
        if (stop - start < KID_KSIZE) { // :755 - the read vanishes (also covers len <= 30)
            if (p.out_taxon && lane == 0) p.out_taxon[r] = -1;
            continue;
        }

        // ------------------------------------------------- 2..4 windows over the trimmed span
        const uintptr_t addr0 = reinterpret_cast<uintptr_t>(p.seq) + g0 + (uint64_t)start;
        const uintptr_t abase = addr0 & ~(uintptr_t)15;
        const int delta = (int)(addr0 - abase);   // staged index of base `start`
        const int nk = stop - start - (KID_KSIZE - 2); // number of k-mer start positions
        const int staged_len = delta + (stop - start + 1);
        uint32_t fin = 0;

        for (int wbase = 0; wbase < nk; wbase += kWindowStarts) {
            __syncwarp();
            { // STAGE
                uint4 v = make_uint4(0, 0, 0, 0);
                if (wbase + 16 * lane < staged_len)
                    v = __ldg(reinterpret_cast<const uint4 *>(abase + (uintptr_t)wbase) + lane);
                uint32_t c0, c1, c2, c3, v0, v1, v2, v3;
                pack4(v.x, p.accept_u, c0, v0);
                pack4(v.y, p.accept_u, c1, v1);
                pack4(v.z, p.accept_u, c2, v2);
                pack4(v.w, p.accept_u, c3, v3);
                strip.codes[lane] = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
                const uint32_t v16 = (v0 << 12) | (v1 << 8) | (v2 << 4) | v3;
                const uint32_t nb = __shfl_down_sync(full, v16, 1);
                if ((lane & 1) == 0) strip.valid[lane >> 1] = (v16 << 16) | nb;
            }
            __syncwarp();

            const int wcount = min(kWindowStarts, nk - wbase);
            for (int c = 0; c < wcount; c += 32 * kUnroll) {
                uint64_t h[kUnroll];
                uint64_t e[kUnroll][4];
                bool act[kUnroll];
                // keys + issue all bucket loads
#pragma unroll
                for (int u = 0; u < kUnroll; u++) {
                    const int j = c + 32 * u + lane; // k-mer start within this window
                    act[u] = false;
                    h[u] = 0;
                    if (j < wcount) {
                        const int t = delta + j;
                        const int w = t >> 4, sh = (t & 15) * 2;
                        const uint32_t w0 = strip.codes[w], w1 = strip.codes[w + 1],
                                       w2 = strip.codes[w + 2];
                        const uint32_t hi = __funnelshift_l(w1, w0, sh);
                        const uint32_t lo = __funnelshift_l(w2, w1, sh);
                        const uint64_t kf = (((uint64_t)hi << 32) | lo) >> 4;
                        const int vw = t >> 5, vs = t & 31;
                        const uint32_t vwin =
                            __funnelshift_l(strip.valid[vw + 1], strip.valid[vw], vs);
                        if ((vwin >> 2) == 0x3FFFFFFFu) {
                            const uint64_t kr = kid_revcomp60(kf);
                            const uint64_t key = kf < kr ? kf : kr; // :528
                            h[u] = kid_hash60(key);
                            act[u] = true;
                        }
                    }
                    if (act[u]) {
                        const uint64_t b = (h[u] >> p.table.rem_bits) & p.table.bucket_mask;
                        kid_load_bucket(p.table.slots + 4 * b, e[u]);
                    }
                }
                // consume in position order
#pragma unroll
                for (int u = 0; u < kUnroll; u++) {
                    if (c + 32 * u >= wcount) break; // warp-uniform
                    uint32_t taxon = 0;
                    if (act[u]) {
                        const uint64_t rem = h[u] & ((1ULL << p.table.rem_bits) - 1ULL);
                        const uint64_t home = (h[u] >> p.table.rem_bits) & p.table.bucket_mask;
                        int j = 0;
                        uint64_t slot = 0;
                        const int res = kid_match_bucket(e[u], rem << KID_DISP_BITS, taxon, j);
                        if (res > 0) slot = 4 * home + (uint64_t)j;
                        else if (res < 0) taxon = kid_lookup_from(p.table, h[u], 1, slot);
                        if (taxon > 1) { // :596-603
                            const uint32_t bit = 1u << (slot & 31);
                            uint32_t *wp = p.seen + (slot >> 5);
                            if (!(*reinterpret_cast<volatile uint32_t *>(wp) & bit)) atomicOr(wp, bit);
                        }
                    }
                    n_lookups += __popc(__ballot_sync(full, act[u]));
                    unsigned m = __ballot_sync(full, taxon > 0);
                    n_hits += __popc(m);
                    while (m) { // ordered left fold over the hits of this chunk (:588-595)
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        const uint32_t tj = __shfl_sync(full, taxon, src);
                        if (fin > 0) { if (tj != fin) fin = kid_msca(p.tree, tj, fin); }
                        else fin = tj;
                    }
                }
            }
        }

        // ---------------------------------------------------------------- 5. COUNT (:613)
        if (lane == 0) {
            if (p.out_taxon) p.out_taxon[r] = (int32_t)fin;
            if (SMEM_HIST) atomicAdd(&hist[fin], 1);
            else atomicAdd(&p.gcount[fin], 1);
        }
    }

    if (lane == 0) {
        if (n_lookups) atomicAdd(p.counters + 0, n_lookups);
        if (n_hits) atomicAdd(p.counters + 1, n_hits);
    }
    if (SMEM_HIST) {
        __syncthreads();
        for (int i = threadIdx.x; i < p.tree.n_taxa; i += blockDim.x) {
            const int v = hist[i];
            if (v) atomicAdd(&p.gcount[i], v);
        }
    }
}

template <bool Q, bool H>
cudaError_t launch_one(const KidClassifyParams &p, int sm_count, cudaStream_t stream)
{
    const size_t smem = sizeof(WarpStrip) * kWarpsPerBlock + (H ? (size_t)p.tree.n_taxa * 4 : 0);
    auto kern = kid_classify_kernel<Q, H>;
    cudaError_t err = cudaSuccess;
    if (smem > 48 * 1024) {
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
    }
    int per_sm = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, KID_CLASSIFY1_THREADS, smem);
    if (err != cudaSuccess) return err;
    if (per_sm < 1) per_sm = 1;
    // persistent grid: a whole number of resident waves, but never more warps than reads
    size_t blocks = (size_t)sm_count * per_sm;
    const size_t need = (p.n_reads + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks > need) blocks = need;
    if (blocks == 0) return cudaSuccess;
    kern<<<(unsigned)blocks, KID_CLASSIFY1_THREADS, smem, stream>>>(p);
    KID_COUNT_LAUNCH();
    return cudaGetLastError();
}

} // namespace

cudaError_t kid_launch_classify(const KidClassifyParams &p, int sm_count, cudaStream_t stream)
{
    const bool hist = (size_t)p.tree.n_taxa * 4 <= KID_SMEM_HIST_MAX_BYTES;
    if (p.qual) return hist ? launch_one<true, true>(p, sm_count, stream)
                            : launch_one<true, false>(p, sm_count, stream);
    return hist ? launch_one<false, true>(p, sm_count, stream)
                : launch_one<false, false>(p, sm_count, stream);
}
