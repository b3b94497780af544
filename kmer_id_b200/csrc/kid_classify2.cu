// kid_classify2.cu - the per-read hot path over the minimizer-addressed table (kid_table2.cuh).
//
// Replaces, for a whole batch of reads, process_qual (newkmer_10nx.cpp:714-760), process_read
// (:452-617), Hashtable::getHash (:204-233) and Tree1::msca (:118-144).
//
// Persistent grid, one 1024-thread block per SM (kid_kernels.cuh); a warp works on one read at a
// time and lane j of chunk c owns the k-mer that starts at base 32c+j of that read.
//   GROUP   a warp takes kGroup (3) consecutive reads at a time; when they fit one 512-base window
//           (always for 150-bp reads) they are loaded, packed and masked ONCE, which cuts the
//           per-read prologue by ~2/3.  Longer reads take the one-read-at-a-time path (windows of
//           448 k-mer starts).
//   LOAD    as soon as the offsets are known the warp issues, together, the coalesced 128-bit
//           loads of the bases (streaming: L1::no_allocate) and 32-byte loads of every read's
//           first/last quality bytes.
//   TRIM    the four `while` loops of process_qual (:727-753) are "first/last position with a
//           property" searches; the common case is answered from the two preloaded quality
//           registers with ballots and shuffles, the rest by a 32-positions-per-step scan.
//   STAGE   each lane packs its 16 bases into 2-bit codes + an ACGT-validity mask in a per-warp
//           shared-memory strip (512 bases per window).
//   KEYS    three LDS + two funnel shifts give the 64 bits that start at the lane's base: the top
//           32 are its 16-mer (minimizer candidate), the top 60 its forward k-mer (:481-517);
//           __brevll gives the reverse complement, min() the canonical key (:528).
//   MINIM   sliding minimum of the 16-mer hashes over 15 positions with 4 shuffle rounds
//           (1,2,4,7) across kUnroll (2) chunks + a 14-lane halo: M(key) without looking at the key.
//   LOOKUP  sector = 4*line(M) + sector(key); kUnroll chunks (64 k-mers) request their 32-byte
//           sectors before the first is consumed.  Lanes that share a minimizer share a line, so
//           a warp-wide load touches ~5 lines instead of 32.
//   FOLD    hits are rare; a ballot finds them and the warp folds them strictly in position order
//           with kid_msca (the fold is order dependent).
//   COUNT   seen bit (atomicOr on the per-sample bitmap) for hits with taxon > 1 (:596-603),
//           gcount[final]++ (:613) in a shared-memory histogram flushed once per block.
#include "kid_kernels.cuh"
#include "kid_readprep.cuh"

#include <cstdlib>
#include <type_traits>

namespace {

constexpr int kWarpsPerBlock = KID_CLASSIFY_THREADS / 32;
constexpr int kWindowStarts = 448; // k-mer starts per staged window: 15 + 447 + 14 + 15 < 512
constexpr int kCodeWords = 44; // 32 + zero padding so that halo lanes never need a bounds check
constexpr int kValidWords = 20; // 16 + zero padding

struct WarpStrip {
    uint32_t codes[kCodeWords];
    uint32_t valid[kValidWords];
    uint32_t kmask[kValidWords];
    int meta[12]; // grouped path: {start, stop, offset in the strip} of each read of the group
};

constexpr int kGroup = 3;          // consecutive reads a warp stages together
constexpr int kGroupMaxSpan = 496; // delta + bytes of the whole group must stay inside the window

// ---- STAGE one 512-base window: 2-bit codes, ACGT validity bits and the "a run of 30 valid bases
// starts here" mask (bit 31 - t%32 of word t/32; the read's own range is applied by the caller)
__device__ __forceinline__ void stage_window(WarpStrip &strip, const uint4 &v, bool accept_u, int lane)
{
    const unsigned full = 0xFFFFFFFFu;
    __syncwarp();
    uint32_t c0, c1, c2, c3, v0, v1, v2, v3;
    pack4(v.x, accept_u, c0, v0);
    pack4(v.y, accept_u, c1, v1);
    pack4(v.z, accept_u, c2, v2);
    pack4(v.w, accept_u, c3, v3);
    strip.codes[lane] = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
    const uint32_t v16 = (v0 << 12) | (v1 << 8) | (v2 << 4) | v3;
    const uint32_t nb = __shfl_down_sync(full, v16, 1);
    if ((lane & 1) == 0) strip.valid[lane >> 1] = (v16 << 16) | nb;
    __syncwarp();
    const int l16 = lane & 15;
    uint64_t x = ((uint64_t)strip.valid[l16] << 32) | strip.valid[l16 + 1];
    x &= x << 1; x &= x << 2; x &= x << 4; x &= x << 8; // runs of 16 valid bases
    x &= x << 14;                                       // runs of 30 (cpos == KSIZE, :526)
    if (lane < 16) strip.kmask[lane] = (uint32_t)(x >> 32);
    __syncwarp();
}

// ---- KEYS .. FOLD over the k-mers that start at window positions [first, last]; the k-mer that
// starts at window position j has its first base at staged index tbase + j
template <int kUnroll>
__device__ __forceinline__ void scan_kmers(const KidClassifyParams &p, const Kid2TableView &tab, const WarpStrip &strip,
                                           int tbase, int first, int last, int lane, uint32_t &fin,
                                           unsigned &lane_lookups, unsigned long long &n_hits)
{
    const unsigned full = 0xFFFFFFFFu;
    auto blockN = [&](int c, auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        // chunks (of 32 k-mers) this iteration has; FULL == all of them, known at compile time
        const int nch = FULL ? kUnroll : min(kUnroll, (last - c + 32) >> 5);
        uint32_t cm[kUnroll + 1]; // 16-mer hashes -> sliding minima
        uint64_t key[kUnroll];
        // KEYS
#pragma unroll
        for (int u = 0; u <= kUnroll; u++) {
            cm[u] = 0xFFFFFFFFu;
            if (FULL || u <= nch) { // chunk nch is the 14-lane halo of chunk nch-1
                const int t = tbase + c + 32 * u + lane; // staged index of this lane's base
                const int w = t >> 4, sh = (t & 15) * 2;
                const uint32_t w0 = strip.codes[w], w1 = strip.codes[w + 1];
                const uint32_t hi = __funnelshift_l(w1, w0, sh);
                const uint32_t rc = kid_rc16(hi);
                cm[u] = kid_mm_hash_canon(min(hi, rc));
                if (u < kUnroll) {
                    // forward key = first 30 of the 32 bases at this position; the reverse complement
                    // of 32 bases is rc16(low half) : rc16(high half) and its low 60 bits are the
                    // reverse complement of the first 30 bases
                    const uint32_t lo = __funnelshift_l(strip.codes[w + 2], w1, sh);
                    const uint64_t kf = (((uint64_t)hi << 32) | lo) >> 4;
                    const uint64_t kr = (((uint64_t)kid_rc16(lo) << 32) | rc) & KID_MASK60;
                    key[u] = kf < kr ? kf : kr; // :528
                }
            }
        }
        // MINIM: window minimum over 15 consecutive positions (1 + 2 + 4 + 7 doubling)
#pragma unroll
        for (int step = 0; step < 4; step++) {
            const int d = step == 0 ? 1 : step == 1 ? 2 : step == 2 ? 4 : 7;
            const int src = lane + d; // shfl takes the source lane modulo 32
            const bool wrap = lane + d >= 32;
            uint32_t s[kUnroll + 1];
#pragma unroll
            for (int u = 0; u <= kUnroll; u++)
                if (FULL || u <= nch) s[u] = __shfl_sync(full, cm[u], src);
#pragma unroll
            for (int u = 0; u < kUnroll; u++)
                if (FULL || u < nch) cm[u] = min(cm[u], wrap ? s[u + 1] : s[u]);
            if (FULL) cm[kUnroll] = min(cm[kUnroll], wrap ? 0xFFFFFFFFu : s[kUnroll]);
            else {
#pragma unroll
                for (int u = 1; u < kUnroll; u++)
                    if (u == nch) cm[u] = min(cm[u], wrap ? 0xFFFFFFFFu : s[u]);
            }
        }
        // LOOKUP: issue every sector load before consuming any.  Inactive lanes (k-mer with an N, or
        // outside the read) read sector 0 instead of branching; their result is ignored.
        uint4 ea[kUnroll], eb[kUnroll];
        uint32_t sec[kUnroll]; // sector index (< 2^32: at most 2^30 lines)
        bool act[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            act[u] = false;
            if (FULL || u < nch) {
                const int j = c + 32 * u + lane, t = tbase + j;
                act[u] = (int32_t)(strip.kmask[t >> 5] << (t & 31)) < 0 && (unsigned)(j - first) <= (unsigned)(last - first);
                lane_lookups += act[u]; // each is one getHash call (:529); summed over lanes at the end
                const uint32_t grp = (cm[u] * 0x9E3779B1u) >> tab.line_shift;
                sec[u] = (grp << tab.sub_bits) | (kid_key_hash32(key[u]) >> (32 - tab.sub_bits));
                kid2_load_sector(tab.sectors + 2 * (uint64_t)(act[u] ? sec[u] : 0u), ea[u], eb[u]);
            }
        }
        // MATCH round 1: just "hit" and "sector full without a match" per lane
        bool hit[kUnroll];
        uint32_t again = 0;
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            hit[u] = false;
            if (FULL || u < nch) {
                const uint32_t klo = (uint32_t)key[u], khi = (uint32_t)(key[u] >> 32) | 0x80000000u;
                const bool h = (ea[u].x == klo && ea[u].y == khi) || (ea[u].z == klo && ea[u].w == khi) ||
                               (eb[u].x == klo && eb[u].y == khi);
                hit[u] = act[u] && h;
                // all three entries carry bit 63 and none matched: the key may live further on
                const bool more = act[u] && !h && (int32_t)(ea[u].y & ea[u].w & eb[u].y) < 0;
                again |= more ? (1u << u) : 0u;
            }
        }
        // MATCH round 2 for the few lanes that met a full sector: all loads first
        if (__any_sync(full, again != 0)) {
#pragma unroll
            for (int u = 0; u < kUnroll; u++)
                if (FULL || u < nch)
                    kid2_load_sector_if(tab.sectors + 2 * ((uint64_t)sec[u] + 1), ea[u], eb[u], (again >> u) & 1u);
#pragma unroll
            for (int u = 0; u < kUnroll; u++) {
                if ((FULL || u < nch) && ((again >> u) & 1u)) {
                    const uint32_t klo = (uint32_t)key[u], khi = (uint32_t)(key[u] >> 32) | 0x80000000u;
                    uint32_t tx = 0;
                    int j = 0;
                    const int res = kid2_match(ea[u], eb[u], klo, khi, tx, j);
                    if (res > 0) {
                        hit[u] = true;
                        sec[u] = sec[u] + 1; // slack sectors follow the last home sector
                    } else if (res < 0) { // rare: third sector and on
                        uint64_t slot = 0;
                        if (kid2_lookup_from(tab, sec[u], key[u], 2, slot)) {
                            hit[u] = true;
                            sec[u] = (uint32_t)(slot / KID2_SLOTS_PER_SECTOR);
                            kid2_load_sector(tab.sectors + 2 * (uint64_t)sec[u], ea[u], eb[u]);
                        }
                    }
                }
            }
        }
        // SEEN + FOLD, strictly in position order; taxa are only extracted when a chunk has hits
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            if (!FULL && u >= nch) break; // warp-uniform
            unsigned m = __ballot_sync(full, hit[u]);
            if (m) {
                uint32_t taxon = 0;
                if (hit[u]) {
                    const uint32_t klo = (uint32_t)key[u], khi = (uint32_t)(key[u] >> 32) | 0x80000000u;
                    int j = 0;
                    kid2_match(ea[u], eb[u], klo, khi, taxon, j);
                    if (taxon > 1) { // :596-603 - fire and forget, the OR is idempotent
                        const uint64_t slot = KID2_SLOTS_PER_SECTOR * (uint64_t)sec[u] + (uint64_t)j;
                        atomicOr(p.seen + (slot >> 5), 1u << (slot & 31));
                    }
                }
                n_hits += __popc(m);
                do { // ordered left fold over the hits of this chunk (:588-595)
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t tj = __shfl_sync(full, taxon, src);
                    if (fin > 0) { if (tj != fin) fin = kid_msca(p.tree, tj, fin); }
                    else fin = tj;
                } while (m);
            }
        }
    };
    for (int c = (first / 32) * 32; c <= last; c += 32 * kUnroll) {
        if (last - c >= 32 * (kUnroll - 1)) blockN(c, std::true_type{});
        else blockN(c, std::false_type{});
    }
}

// kUnroll = chunks of 32 k-mers whose sector loads are in flight together; kMinBlocks = resident
// blocks per SM the register budget is cut for (launch bounds)
template <bool HAS_QUAL, bool SMEM_HIST, int kUnroll, int kMinBlocks>
__global__ void __launch_bounds__(KID_CLASSIFY_THREADS, kMinBlocks)
kid_classify2_kernel(const KidClassifyParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpStrip *strips = reinterpret_cast<WarpStrip *>(smem_raw);
    int *hist = reinterpret_cast<int *>(smem_raw + sizeof(WarpStrip) * kWarpsPerBlock);

    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    WarpStrip &strip = strips[warp_in_block];
    const Kid2TableView tab = p.table2;

    if (SMEM_HIST) {
        for (int i = threadIdx.x; i < p.tree.n_taxa; i += blockDim.x) hist[i] = 0;
        __syncthreads();
    }
    if (lane < kCodeWords - 32) strip.codes[32 + lane] = 0;
    if (lane < kValidWords - 16) { strip.valid[16 + lane] = 0; strip.kmask[16 + lane] = 0; }

    unsigned lane_lookups = 0;     // per lane (< 2^32 per launch), reduced once at the end
    unsigned long long n_hits = 0; // warp-uniform

    // the read's final taxon: per-read output and gcount[final]++ (:613)
    auto finish_read = [&](size_t r, int start, int stop, bool kept, uint32_t fin) {
        if (lane == 0) {
            if (p.out_span) {
                p.out_span[2 * r] = (uint32_t)start;
                p.out_span[2 * r + 1] = (uint32_t)stop;
            }
            if (p.out_taxon) p.out_taxon[r] = kept ? (int32_t)fin : -1;
            if (kept) {
                if (SMEM_HIST) atomicAdd(&hist[fin], 1);
                else atomicAdd(&p.gcount[fin], 1);
            }
        }
    };

    const size_t n_groups = (p.n_reads + kGroup - 1) / kGroup;
    const size_t warps_total = (size_t)gridDim.x * kWarpsPerBlock;
    for (size_t grp = (size_t)blockIdx.x * kWarpsPerBlock + warp_in_block; grp < n_groups; grp += warps_total) {
        const size_t r0 = grp * kGroup;
        const int nr = (int)min((size_t)kGroup, p.n_reads - r0);
        // offsets of the group's reads relative to its first read
        const uint64_t g0 = __ldg(p.off + r0) - p.off_bias;
        int rel[kGroup + 1]; // saturated: anything beyond the window size only has to fail the test below
        rel[0] = 0;
#pragma unroll
        for (int i = 1; i <= kGroup; i++) {
            const uint64_t d = i <= nr ? __ldg(p.off + r0 + i) - p.off_bias - g0 : (uint64_t)rel[i - 1];
            rel[i] = d > 0x7FFFFFFFull ? 0x7FFFFFFF : (int)d;
        }
        const uintptr_t addr0 = reinterpret_cast<uintptr_t>(p.seq) + g0;
        const uintptr_t abase = addr0 & ~(uintptr_t)15;
        const int delta = (int)(addr0 - abase); // staged index of the group's first base

        if (rel[kGroup] <= kGroupMaxSpan - delta) {
            // ---- grouped path: one load / pack / mask for all reads of the group
            uint4 v = make_uint4(0, 0, 0, 0);
            if (16 * lane < delta + rel[kGroup]) v = load_stream16(reinterpret_cast<const uint4 *>(abase) + lane);
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
                for (int i = 0; i < kGroup; i++) {
                    int st, sp;
                    trim_read(q + rel[i], rel[i + 1] - rel[i], qa[i], qb[i], lane, st, sp);
                    if (lane == 0) { strip.meta[3 * i] = st; strip.meta[3 * i + 1] = sp; strip.meta[3 * i + 2] = delta + rel[i]; }
                }
            } else if (lane == 0) {
#pragma unroll
                for (int i = 0; i < kGroup; i++) {
                    strip.meta[3 * i] = 0;
                    strip.meta[3 * i + 1] = rel[i + 1] - rel[i] - 1;
                    strip.meta[3 * i + 2] = delta + rel[i];
                }
            }
            stage_window(strip, v, p.accept_u, lane); // (its __syncwarp()s also publish meta[])
            for (int i = 0; i < nr; i++) {
                const int start = strip.meta[3 * i], stop = strip.meta[3 * i + 1], tbase = strip.meta[3 * i + 2];
                const bool kept = stop - start >= KID_KSIZE; // :755
                uint32_t fin = 0;
                if (kept)
                    scan_kmers<kUnroll>(p, tab, strip, tbase, start, stop - (KID_KSIZE - 1), lane, fin, lane_lookups, n_hits);
                finish_read(r0 + i, start, stop, kept, fin);
            }
            __syncwarp(); // meta[] is rewritten by the next group
            continue;
        }

        // ---- one read at a time (long reads): windows of kWindowStarts k-mer starts
        for (int i = 0; i < nr; i++) {
            const uint64_t gi = __ldg(p.off + r0 + i) - p.off_bias; // exact 64-bit offsets (one read < 2^31 bases)
            const int len = (int)(__ldg(p.off + r0 + i + 1) - p.off_bias - gi);
            const uintptr_t a0 = reinterpret_cast<uintptr_t>(p.seq) + gi;
            const uintptr_t ab = a0 & ~(uintptr_t)15;
            const int dl = (int)(a0 - ab);
            int start = 0, stop = len - 1;
            if (HAS_QUAL && len > 0) {
                const signed char *q = reinterpret_cast<const signed char *>(p.qual) + gi;
                const int qa = lane < len ? (int)q[lane] : -128;
                const int qb = lane < len ? (int)q[len - 1 - lane] : -128;
                trim_read(q, len, qa, qb, lane, start, stop);
            }
            const bool kept = stop - start >= KID_KSIZE;
            uint32_t fin = 0;
            if (kept) {
                const int last_start = stop - (KID_KSIZE - 1);
                for (int wb = 0; wb <= last_start; wb += kWindowStarts) {
                    if (wb + kWindowStarts <= start) continue; // window entirely before the trimmed span
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (wb + 16 * lane < dl + len) v = load_stream16(reinterpret_cast<const uint4 *>(ab + (uintptr_t)wb) + lane);
                    stage_window(strip, v, p.accept_u, lane);
                    scan_kmers<kUnroll>(p, tab, strip, dl, max(0, start - wb), min(kWindowStarts - 1, last_start - wb),
                                        lane, fin, lane_lookups, n_hits);
                }
            }
            finish_read(r0 + i, start, stop, kept, fin);
        }
    }

    unsigned long long n_lookups = lane_lookups;
    for (int o = 16; o; o >>= 1) n_lookups += __shfl_xor_sync(0xFFFFFFFFu, n_lookups, o);
    if (lane == 0) {
        if (n_lookups) atomicAdd(p.counters + 0, n_lookups);
        if (n_hits) atomicAdd(p.counters + 1, n_hits);
    }
    if (SMEM_HIST) {
        __syncthreads();
        for (int i = threadIdx.x; i < p.tree.n_taxa; i += blockDim.x) {
            const int cnt = hist[i];
            if (cnt) atomicAdd(&p.gcount[i], cnt);
        }
    }
}

template <bool Q, bool H, int U, int MB>
cudaError_t launch_one(const KidClassifyParams &p, int sm_count, cudaStream_t stream)
{
    const size_t smem = sizeof(WarpStrip) * kWarpsPerBlock + (H ? (size_t)p.tree.n_taxa * 4 : 0);
    auto kern = kid_classify2_kernel<Q, H, U, MB>;
    cudaError_t err = cudaSuccess;
    if (smem > 48 * 1024) {
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
    }
    int per_sm = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, KID_CLASSIFY_THREADS, smem);
    if (err != cudaSuccess) return err;
    if (per_sm < 1) per_sm = 1;
    // shared memory and L1 are one 256 KB array: reserve only what the resident blocks use, the
    // rest caches the read stream, the taxonomy rows and the spill slots
    if (getenv("KID_NO_CARVEOUT") == nullptr) {
        const size_t need = (size_t)per_sm * (smem + 1024);
        int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
        if (pct > 100) pct = 100;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct); // a hint: failure is harmless
    }
    size_t blocks = (size_t)sm_count * per_sm; // persistent: a whole number of resident waves
    const size_t need = ((p.n_reads + kGroup - 1) / kGroup + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks > need) blocks = need;
    if (blocks == 0) return cudaSuccess;
    kern<<<(unsigned)blocks, KID_CLASSIFY_THREADS, smem, stream>>>(p);
    KID_COUNT_LAUNCH();
    return cudaGetLastError();
}

template <int U, int MB>
cudaError_t launch_variant(const KidClassifyParams &p, int sm_count, cudaStream_t stream)
{
    const bool hist = (size_t)p.tree.n_taxa * 4 <= KID_SMEM_HIST_MAX_BYTES;
    if (p.qual) return hist ? launch_one<true, true, U, MB>(p, sm_count, stream)
                            : launch_one<true, false, U, MB>(p, sm_count, stream);
    return hist ? launch_one<false, true, U, MB>(p, sm_count, stream)
                : launch_one<false, false, U, MB>(p, sm_count, stream);
}

} // namespace

cudaError_t kid_launch_classify2(const KidClassifyParams &p, int sm_count, cudaStream_t stream)
{
#ifdef KID_TUNE_VARIANTS // experiment builds only: pick the variant with KID_TUNE=<n>
    static const int tune = getenv("KID_TUNE") ? atoi(getenv("KID_TUNE")) : 0;
    switch (tune) {
    case 1: return launch_variant<1, 1024 / KID_CLASSIFY_THREADS>(p, sm_count, stream);
    case 2: return launch_variant<3, 1024 / KID_CLASSIFY_THREADS>(p, sm_count, stream);
    case 3: return launch_variant<4, 1024 / KID_CLASSIFY_THREADS>(p, sm_count, stream);
    default: break;
    }
#endif
    // measured on B200 (tools/run_tune.sh): 2 chunks in flight at 32 warps/SM (64 registers) beats
    // 4 chunks at 24 warps/SM by 27 % - the kernel is issue/latency bound, not DRAM bound
#ifndef KID_CLASSIFY_MINBLOCKS
#define KID_CLASSIFY_MINBLOCKS (1024 / KID_CLASSIFY_THREADS) // 32 warps per SM at 64 registers per thread
#endif
    return launch_variant<2, KID_CLASSIFY_MINBLOCKS>(p, sm_count, stream);
}
