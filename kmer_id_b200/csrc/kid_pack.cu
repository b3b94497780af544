// kid_pack.cu - raw read bytes on the device -> a packed read batch (kid_kernels.cuh).
//
// Replaces process_qual (newkmer_10nx.cpp:714-760) and the per-base switch of process_read
// (:477-525) for a whole batch: trim each read by its qualities, test every base of the trimmed
// span for ACGTacgt (+Uu), and write the span as 2-bit codes that start on a word boundary, plus
// validity words.  A streaming kernel; the k-mer scan (kid_classify3.cu) then never touches text or
// qualities.
//
// A warp takes 32 consecutive reads at a time.
//   TRIM  one lane per read: a read whose first and last quality byte and 4-base window pass (nearly
//         all good reads) is decided from 8 bytes; a short bad head or tail is walked by the lane itself
//         (process_qual's loops as they stand, up to 96 steps); only longer runs go through the
//         warp-cooperative searches of kid_readprep.cuh, one read after the other.
//   PACK  one lane per 32 bases of trimmed sequence, over all 32 reads (a prefix sum of the reads'
//         unit counts in shared memory, a 5-step binary search per unit): nine aligned 32-bit loads,
//         byte realignment with funnel shifts, SIMD-in-word packing (kid_readprep.cuh) -> two code
//         words and one validity word.  Reads of any length take the same path.
#include "kid_kernels.cuh"
#include "kid_readprep.cuh"

namespace {

constexpr int kPackThreads = 256;

struct PackTile { // one warp's 32 reads
    uint64_t off[32];  // first base of the read, relative to the batch
    uint32_t pre[33];  // units (32 bases) before read i
    int start[32];
    uint32_t tlen[32]; // 0 for a read the length rule drops
    uint32_t flag[32]; // KID_PK_FLAG once a base outside ACGT was seen
};

// 16 validity bits (first base in bit 15) -> 32-bit mask with both bits of every valid base set
__device__ __forceinline__ uint32_t spread16(uint32_t v)
{
    v = (v | (v << 8)) & 0x00FF00FFu;
    v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v | (v << 1);
}

// 16 bases in four little-endian words -> code word + 16 validity bits (first base on top)
__device__ __forceinline__ void pack16(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, bool accept_u,
                                       uint32_t &code, uint32_t &v16)
{
    uint32_t c0, c1, c2, c3, v0, v1, v2, v3;
    pack4(x0, accept_u, c0, v0);
    pack4(x1, accept_u, c1, v1);
    pack4(x2, accept_u, c2, v2);
    pack4(x3, accept_u, c3, v3);
    code = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
    v16 = (v0 << 12) | (v1 << 8) | (v2 << 4) | v3;
}

template <bool HAS_QUAL>
__global__ void __launch_bounds__(kPackThreads)
kid_pack_kernel(const KidPackParams p)
{
    __shared__ PackTile tiles[kPackThreads / 32];
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    PackTile &t = tiles[threadIdx.x >> 5];

    const size_t n_tiles = (p.n_reads + 31) / 32;
    const size_t warps_total = (size_t)gridDim.x * (kPackThreads / 32);
    for (size_t tile = (size_t)blockIdx.x * (kPackThreads / 32) + (threadIdx.x >> 5); tile < n_tiles; tile += warps_total) {
        const size_t r = tile * 32 + (size_t)lane;
        const bool have = r < p.n_reads;
        uint64_t o = 0;
        int len = 0;
        if (have) {
            o = __ldg(p.off + r) - p.off_bias;
            const uint64_t l64 = __ldg(p.off + r + 1) - p.off_bias - o;
            len = l64 > 0x7FFFFFFFull ? 0x7FFFFFFF : (int)l64; // one read < 2^31 bases
        }
        // ---- TRIM (:724-753)
        int start = 0, stop = len - 1;
        if (HAS_QUAL) {
            const signed char *q = reinterpret_cast<const signed char *>(p.qual) + o;
            bool slow = have && len > 0;
            if (have && len >= 6) {
                const int a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3];
                const int b0 = q[len - 1], b1 = q[len - 2], b2 = q[len - 3], b3 = q[len - 4];
                // both end bases and both end windows pass: nothing to trim
                slow = !(a0 >= 49 && b0 >= 49 && a0 + a1 + a2 + a3 - 128 >= 68 && b0 + b1 + b2 + b3 - 128 >= 68);
            }
            // a short bad head or tail: the lane walks it itself, all such lanes of the warp at once; only a
            // long run of bad qualities goes on to the warp-cooperative searches
            if (slow) slow = !trim_read_lane(q, len, 96, start, stop);
            unsigned todo = __ballot_sync(full, slow);
            while (todo) { // the warp-cooperative search, one read after the other
                const int i = __ffs(todo) - 1;
                todo &= todo - 1;
                const uint64_t oi = __shfl_sync(full, o, i);
                const int li = __shfl_sync(full, len, i);
                const signed char *qi = reinterpret_cast<const signed char *>(p.qual) + oi;
                const int qa = lane < li ? (int)qi[lane] : -128;
                const int qb = lane < li ? (int)qi[li - 1 - lane] : -128;
                int st, sp;
                trim_read(qi, li, qa, qb, lane, st, sp);
                if (lane == i) { start = st; stop = sp; }
            }
        }
        const int tl = stop - start + 1;
        const bool kept = have && tl > KID_KSIZE; // :755
        const uint32_t units = kept ? ((uint32_t)tl + 31u) >> 5 : 0u;
        uint32_t incl = units; // inclusive prefix sum over the warp
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(full, incl, d);
            if (lane >= d) incl += v;
        }
        __syncwarp();
        t.off[lane] = o;
        t.start[lane] = start;
        t.tlen[lane] = kept ? (uint32_t)tl : 0u;
        t.flag[lane] = 0;
        t.pre[lane + 1] = incl;
        if (lane == 0) t.pre[0] = 0;
        __syncwarp();
        const uint32_t total = t.pre[32];
        // ---- PACK: one lane per unit of 32 bases
        for (uint32_t w = lane; w < total; w += 32) {
            int ri = 0; // the read this unit belongs to: the last one with pre[ri] <= w
#pragma unroll
            for (int step = 16; step; step >>= 1)
                if (t.pre[ri + step] <= w) ri += step;
            const uint32_t u = w - t.pre[ri];
            const uint32_t tlr = t.tlen[ri];
            const int nb = min(32, (int)(tlr - 32u * u)); // bases of this unit, >= 1
            const uint64_t ro = t.off[ri];
            const uintptr_t src = reinterpret_cast<uintptr_t>(p.seq) + ro + (uint64_t)t.start[ri] + 32ull * u;
            const uint32_t *al = reinterpret_cast<const uint32_t *>(src & ~(uintptr_t)3);
            const int bsh = (int)(src & 3) * 8;
            const int last_word = (int)(((src & 3) + (uintptr_t)nb - 1) >> 2); // aligned words that hold a needed byte: 0..last_word
            uint32_t a[9];
#pragma unroll
            for (int k = 0; k < 9; k++) a[k] = k <= last_word ? __ldg(al + k) : 0u;
            uint32_t x[8];
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = __funnelshift_r(a[k], a[k + 1], bsh);
            uint32_t c0, c1, v0, v1;
            pack16(x[0], x[1], x[2], x[3], p.accept_u, c0, v0);
            pack16(x[4], x[5], x[6], x[7], p.accept_u, c1, v1);
            uint32_t valid = (v0 << 16) | v1;
            const uint32_t want = nb >= 32 ? 0xFFFFFFFFu : 0xFFFFFFFFu << (32 - nb);
            valid &= want; // bytes behind the read's last base are not part of it
            if (valid != want) {
                c0 &= spread16(valid >> 16); // other bases: code 0
                c1 &= spread16(valid & 0xFFFFu);
                atomicOr(&t.flag[ri], KID_PK_FLAG);
            }
            if (nb < 16) c0 &= 0xFFFFFFFFu << (2 * (16 - nb));
            if (nb < 32) c1 = nb > 16 ? c1 & (0xFFFFFFFFu << (2 * (32 - nb))) : 0u;
            const uint32_t wf = (uint32_t)kid_pack_word_index(ro, tile * 32 + (size_t)ri);
            const uint32_t cwn = (tlr + 15u) >> 4;
            p.words[wf + 2 * u] = c0;
            if (2 * u + 1 < cwn) p.words[wf + 2 * u + 1] = c1;
            p.words[wf + cwn + u] = valid; // always written (the layout reserves the room); read only if flagged
        }
        __syncwarp();
        if (have) {
            p.meta[r] = make_uint2((uint32_t)kid_pack_word_index(o, r) | t.flag[lane], t.tlen[lane]);
            if (p.out_span) { p.out_span[2 * r] = (uint32_t)start; p.out_span[2 * r + 1] = (uint32_t)stop; }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const uint64_t total = __ldg(p.off + p.n_reads) - p.off_bias;
        p.meta[p.n_reads] = make_uint2((uint32_t)kid_pack_word_index(total, p.n_reads), 0u);
    }
}

// ---- dense batch -> packed batch ----------------------------------------------------------------
// What arrives over PCIe from a host parser is the dense form (no padding, offsets only, non-ACGT bases
// as a position list); the scan wants every read on a word boundary with validity words.  Same shape as
// the PACK phase above: a warp takes 32 reads, one lane per 32 bases; the source is already 2-bit codes,
// so a unit is three words and two funnel shifts.
struct ExpandTile {
    uint32_t b0[32];  // first base of the read, relative to the batch's codes
    uint32_t pre[33]; // units before read i
    uint32_t tlen[32];
};

__global__ void __launch_bounds__(kPackThreads)
kid_expand_kernel(const KidExpandParams p)
{
    __shared__ ExpandTile tiles[kPackThreads / 32];
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    ExpandTile &t = tiles[threadIdx.x >> 5];
    const size_t n_tiles = (p.n_reads + 31) / 32;
    const size_t warps_total = (size_t)gridDim.x * (kPackThreads / 32);
    for (size_t tile = (size_t)blockIdx.x * (kPackThreads / 32) + (threadIdx.x >> 5); tile < n_tiles; tile += warps_total) {
        const size_t r = tile * 32 + (size_t)lane;
        const bool have = r < p.n_reads;
        uint32_t b0 = 0, tl = 0;
        if (have) {
            const uint32_t a = __ldg(p.boff + r);
            tl = __ldg(p.boff + r + 1) - a;
            b0 = a - p.bias;
        }
        const uint32_t fw = __ldg(p.flagbits + tile); // the flags of this tile's 32 reads
        const uint32_t units = (tl + 31u) >> 5;
        uint32_t incl = units;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(full, incl, d);
            if (lane >= d) incl += v;
        }
        __syncwarp();
        t.b0[lane] = b0;
        t.tlen[lane] = tl;
        t.pre[lane + 1] = incl;
        if (lane == 0) t.pre[0] = 0;
        __syncwarp();
        const uint32_t total = t.pre[32];
        for (uint32_t w = lane; w < total; w += 32) {
            int ri = 0;
#pragma unroll
            for (int step = 16; step; step >>= 1)
                if (t.pre[ri + step] <= w) ri += step;
            const uint32_t u = w - t.pre[ri];
            const uint32_t tlr = t.tlen[ri], rb = t.b0[ri];
            const int nb = min(32, (int)(tlr - 32u * u));
            const uint32_t sb = rb + 32u * u; // first base of this unit in the codes
            const uint32_t *src = p.codes + (sb >> 4);
            const int sh = (int)(sb & 15u) * 2;
            const uint32_t a0 = __ldg(src), a1 = __ldg(src + 1), a2 = __ldg(src + 2);
            uint32_t c0 = __funnelshift_l(a1, a0, sh), c1 = __funnelshift_l(a2, a1, sh);
            if (nb < 16) c0 &= 0xFFFFFFFFu << (2 * (16 - nb));
            if (nb < 32) c1 = nb > 16 ? c1 & (0xFFFFFFFFu << (2 * (32 - nb))) : 0u;
            uint32_t valid = nb >= 32 ? 0xFFFFFFFFu : 0xFFFFFFFFu << (32 - nb);
            const bool flagged = (fw >> ri) & 1u;
            if (flagged && p.n_inv) { // clear the bits of the listed positions that fall into this unit
                const uint32_t lo = sb + p.bias, hi = lo + (uint32_t)nb; // stream positions [lo, hi)
                uint32_t a = 0, b = p.n_inv; // first entry >= lo
                while (a < b) {
                    const uint32_t m = (a + b) >> 1;
                    if (__ldg(p.inv + m) < lo) a = m + 1; else b = m;
                }
                for (; a < p.n_inv; a++) {
                    const uint32_t pos = __ldg(p.inv + a);
                    if (pos >= hi) break;
                    valid &= ~(0x80000000u >> (pos - lo));
                }
            }
            const uint32_t wf = (uint32_t)kid_pack_word_index(rb, tile * 32 + (size_t)ri);
            const uint32_t cwn = (tlr + 15u) >> 4;
            p.words[wf + 2 * u] = c0;
            if (2 * u + 1 < cwn) p.words[wf + 2 * u + 1] = c1;
            p.words[wf + cwn + u] = valid;
        }
        if (have) p.meta[r] = make_uint2((uint32_t)kid_pack_word_index(b0, r) | (((fw >> lane) & 1u) ? KID_PK_FLAG : 0u), tl);
        __syncwarp();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const uint32_t total = __ldg(p.boff + p.n_reads) - p.bias;
        p.meta[p.n_reads] = make_uint2((uint32_t)kid_pack_word_index(total, p.n_reads), 0u);
    }
}

} // namespace

cudaError_t kid_launch_pack(const KidPackParams &p, int sm_count, cudaStream_t stream)
{
    if (p.n_reads == 0) return cudaSuccess;
    const size_t warps = (p.n_reads + 31) / 32;
    size_t blocks = (warps + kPackThreads / 32 - 1) / (kPackThreads / 32);
    const size_t cap = (size_t)sm_count * 8; // 2048 threads per SM: 8 resident blocks, grid-stride beyond
    if (blocks > cap) blocks = cap;
    if (p.qual) kid_pack_kernel<true><<<(unsigned)blocks, kPackThreads, 0, stream>>>(p);
    else kid_pack_kernel<false><<<(unsigned)blocks, kPackThreads, 0, stream>>>(p);
    KID_COUNT_LAUNCH();
    return cudaGetLastError();
}

cudaError_t kid_launch_expand(const KidExpandParams &p, int sm_count, cudaStream_t stream)
{
    if (p.n_reads == 0) return cudaSuccess;
    const size_t warps = (p.n_reads + 31) / 32;
    size_t blocks = (warps + kPackThreads / 32 - 1) / (kPackThreads / 32);
    const size_t cap = (size_t)sm_count * 8;
    if (blocks > cap) blocks = cap;
    kid_expand_kernel<<<(unsigned)blocks, kPackThreads, 0, stream>>>(p);
    KID_COUNT_LAUNCH();
    return cudaGetLastError();
}
