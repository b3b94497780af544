// kid_pack.cu - raw read bytes on the device -> a packed read batch (kid_kernels.cuh).
//
// Replaces process_qual (newkmer_10nx.cpp:714-760) and the per-base switch of process_read
// (:477-525) for a whole batch: trim each read by its qualities, test every base of the trimmed
// span for ACGTacgt (+Uu), and write the span as 2-bit codes that start on a word boundary, plus
// validity words when a base fails the test.  A streaming kernel: 2 x L bytes in, ~L/4 bytes out per
// read; the k-mer scan (kid_classify3.cu) then never touches text or qualities.
//
// A warp takes 3 consecutive reads at a time; when they fit one 512-base window (always for 150-bp
// reads) they are loaded with one coalesced 128-bit load per lane, and every read's first / last 32
// quality bytes are requested before any is used.  Longer reads are walked in windows of 480 bases
// and always carry validity words.
#include "kid_kernels.cuh"
#include "kid_readprep.cuh"

namespace {

constexpr int kPackThreads = 256;
constexpr int kCodeWords = 36;  // 32 + zero padding
constexpr int kValidWords = 20; // 16 + zero padding
constexpr int kGroup = 3;
constexpr int kGroupMaxSpan = 496;
constexpr int kLongStep = 480; // bases per window of the long path: 30 code words, 15 validity words

struct PackStrip {
    uint32_t codes[kCodeWords];
    uint32_t valid[kValidWords];
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

__device__ __forceinline__ void stage_bases(PackStrip &strip, const uint4 &v, bool accept_u, int lane)
{
    const unsigned full = 0xFFFFFFFFu;
    __syncwarp();
    uint32_t c0, c1, c2, c3, v0, v1, v2, v3;
    pack4(v.x, accept_u, c0, v0);
    pack4(v.y, accept_u, c1, v1);
    pack4(v.z, accept_u, c2, v2);
    pack4(v.w, accept_u, c3, v3);
    const uint32_t v16 = (v0 << 12) | (v1 << 8) | (v2 << 4) | v3;
    strip.codes[lane] = ((c0 << 24) | (c1 << 16) | (c2 << 8) | c3) & spread16(v16); // other bases: code 0
    const uint32_t nb = __shfl_down_sync(full, v16, 1);
    if ((lane & 1) == 0) strip.valid[lane >> 1] = (v16 << 16) | nb;
    __syncwarp();
}

// words of `n` bases that start at staged index tb: lane k gets code word k (16 bases) and validity
// word k (32 bases), both cut off after base n
__device__ __forceinline__ void extract(const PackStrip &strip, int tb, int n, int lane, uint32_t &code, uint32_t &valid)
{
    {
        const int t = tb + 16 * lane, w = t >> 4, sh = (t & 15) * 2, rem = n - 16 * lane;
        code = 0;
        if (rem > 0 && w + 1 < kCodeWords) {
            code = __funnelshift_l(strip.codes[w + 1], strip.codes[w], sh);
            if (rem < 16) code &= ~0u << (2 * (16 - rem));
        }
    }
    {
        const int t = tb + 32 * lane, w = t >> 5, sh = t & 31, rem = n - 32 * lane;
        valid = 0;
        if (rem > 0 && w + 1 < kValidWords) {
            valid = __funnelshift_l(strip.valid[w + 1], strip.valid[w], sh);
            if (rem < 32) valid &= ~0u << (32 - rem);
        }
    }
}

template <bool HAS_QUAL>
__global__ void __launch_bounds__(kPackThreads)
kid_pack_kernel(const KidPackParams p)
{
    __shared__ PackStrip strips[kPackThreads / 32];
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    PackStrip &strip = strips[threadIdx.x >> 5];
    if (lane < kCodeWords - 32) strip.codes[32 + lane] = 0;
    if (lane < kValidWords - 16) strip.valid[16 + lane] = 0;

    const size_t n_groups = (p.n_reads + kGroup - 1) / kGroup;
    const size_t warps_total = (size_t)gridDim.x * (kPackThreads / 32);
    for (size_t grp = (size_t)blockIdx.x * (kPackThreads / 32) + (threadIdx.x >> 5); grp < n_groups; grp += warps_total) {
        const size_t r0 = grp * kGroup;
        const int nr = (int)min((size_t)kGroup, p.n_reads - r0);
        const uint64_t g0 = __ldg(p.off + r0) - p.off_bias;
        int rel[kGroup + 1]; // saturated: anything beyond the window only has to fail the test below
        rel[0] = 0;
#pragma unroll
        for (int i = 1; i <= kGroup; i++) {
            const uint64_t d = i <= nr ? __ldg(p.off + r0 + i) - p.off_bias - g0 : (uint64_t)rel[i - 1];
            rel[i] = d > 0x7FFFFFFFull ? 0x7FFFFFFF : (int)d;
        }
        const uintptr_t addr0 = reinterpret_cast<uintptr_t>(p.seq) + g0;
        const uintptr_t abase = addr0 & ~(uintptr_t)15;
        const int delta = (int)(addr0 - abase);

        if (rel[kGroup] <= kGroupMaxSpan - delta) {
            // ---- grouped path: one load / pack for all reads of the group
            uint4 v = make_uint4(0, 0, 0, 0);
            if (16 * lane < delta + rel[kGroup]) v = load_stream16(reinterpret_cast<const uint4 *>(abase) + lane);
            int st[kGroup], sp[kGroup];
            if (HAS_QUAL) {
                const signed char *q = reinterpret_cast<const signed char *>(p.qual) + g0;
                int qa[kGroup], qb[kGroup];
#pragma unroll
                for (int i = 0; i < kGroup; i++) { // all quality loads in flight before any is used
                    const int len = rel[i + 1] - rel[i];
                    qa[i] = lane < len ? (int)q[rel[i] + lane] : -128;
                    qb[i] = lane < len ? (int)q[rel[i + 1] - 1 - lane] : -128;
                }
#pragma unroll
                for (int i = 0; i < kGroup; i++) trim_read(q + rel[i], rel[i + 1] - rel[i], qa[i], qb[i], lane, st[i], sp[i]);
            } else {
#pragma unroll
                for (int i = 0; i < kGroup; i++) { st[i] = 0; sp[i] = rel[i + 1] - rel[i] - 1; }
            }
            stage_bases(strip, v, p.accept_u, lane);
#pragma unroll
            for (int i = 0; i < kGroup; i++) {
                if (i >= nr) break;
                const size_t r = r0 + i;
                const int tl = sp[i] - st[i] + 1;
                const bool kept = tl > KID_KSIZE; // :755
                const uint32_t wf = (uint32_t)kid_pack_word_index(g0 + (uint64_t)rel[i], r);
                uint32_t flag = 0;
                if (kept) {
                    uint32_t code, valid;
                    extract(strip, delta + rel[i] + st[i], tl, lane, code, valid);
                    const int cwn = (tl + 15) >> 4, vwn = (tl + 31) >> 5;
                    const int rem = tl - 32 * lane;
                    const uint32_t want = rem >= 32 ? ~0u : (rem > 0 ? ~0u << (32 - rem) : 0u);
                    flag = __any_sync(full, valid != want) ? KID_PK_FLAG : 0u;
                    if (lane < cwn) p.words[wf + lane] = code;
                    if (flag && lane < vwn) p.words[wf + cwn + lane] = valid;
                }
                if (lane == 0) {
                    p.meta[r] = make_uint2(wf | flag, kept ? (uint32_t)tl : 0u);
                    if (p.out_span) { p.out_span[2 * r] = (uint32_t)st[i]; p.out_span[2 * r + 1] = (uint32_t)sp[i]; }
                }
            }
            continue;
        }

        // ---- one read at a time (long reads): windows of kLongStep bases from the trimmed start
        for (int i = 0; i < nr; i++) {
            const size_t r = r0 + i;
            const uint64_t gi = __ldg(p.off + r) - p.off_bias;
            const int len = (int)(__ldg(p.off + r + 1) - p.off_bias - gi);
            int start = 0, stop = len - 1;
            if (HAS_QUAL && len > 0) {
                const signed char *q = reinterpret_cast<const signed char *>(p.qual) + gi;
                const int qa = lane < len ? (int)q[lane] : -128;
                const int qb = lane < len ? (int)q[len - 1 - lane] : -128;
                trim_read(q, len, qa, qb, lane, start, stop);
            }
            const int tl = stop - start + 1;
            const bool kept = tl > KID_KSIZE;
            const uint32_t wf = (uint32_t)kid_pack_word_index(gi, r);
            if (kept) {
                const int cwn = (tl + 15) >> 4;
                for (int wb = 0; wb < tl; wb += kLongStep) {
                    const uintptr_t a0 = reinterpret_cast<uintptr_t>(p.seq) + gi + (uint64_t)start + (uint64_t)wb;
                    const uintptr_t ab = a0 & ~(uintptr_t)15;
                    const int dl = (int)(a0 - ab), nb = min(kLongStep, tl - wb);
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (16 * lane < dl + nb) v = load_stream16(reinterpret_cast<const uint4 *>(ab) + lane);
                    stage_bases(strip, v, p.accept_u, lane);
                    uint32_t code, valid;
                    extract(strip, dl, nb, lane, code, valid);
                    if (16 * lane < nb) p.words[wf + (wb >> 4) + lane] = code;
                    if (32 * lane < nb) p.words[wf + cwn + (wb >> 5) + lane] = valid;
                }
            }
            if (lane == 0) {
                p.meta[r] = make_uint2(wf | (kept ? KID_PK_FLAG : 0u), kept ? (uint32_t)tl : 0u);
                if (p.out_span) { p.out_span[2 * r] = (uint32_t)start; p.out_span[2 * r + 1] = (uint32_t)stop; }
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const uint64_t total = __ldg(p.off + p.n_reads) - p.off_bias;
        p.meta[p.n_reads] = make_uint2((uint32_t)kid_pack_word_index(total, p.n_reads), 0u);
    }
}

} // namespace

cudaError_t kid_launch_pack(const KidPackParams &p, int sm_count, cudaStream_t stream)
{
    if (p.n_reads == 0) return cudaSuccess;
    const size_t warps = (p.n_reads + kGroup - 1) / kGroup;
    size_t blocks = (warps + kPackThreads / 32 - 1) / (kPackThreads / 32);
    const size_t cap = (size_t)sm_count * 8; // 2048 threads per SM: 8 resident blocks, grid-stride beyond
    if (blocks > cap) blocks = cap;
    if (p.qual) kid_pack_kernel<true><<<(unsigned)blocks, kPackThreads, 0, stream>>>(p);
    else kid_pack_kernel<false><<<(unsigned)blocks, kPackThreads, 0, stream>>>(p);
    KID_COUNT_LAUNCH();
    return cudaGetLastError();
}
