// kid_classify3.cu - the per-read hot path over PACKED read batches (kid_kernels.cuh) and the
// minimizer-addressed table (kid_table2.cuh).
//
// Replaces, for a whole batch of reads, process_read (newkmer_10nx.cpp:452-617), Hashtable::getHash
// (:204-233) and Tree1::msca (:118-144).  process_qual's trim (:714-760) and the ACGT test (:480-524)
// happened when the batch was packed (kid_pack_kernel from text, kid_expand_kernel from the dense batch
// a host parser ships, or kid_pack_reads on the host), so a read arrives as `tlen` 2-bit codes that
// start on a word boundary, plus validity words only if it contains a base outside ACGTacgt.
//
// Persistent grid, one 1024-thread block per SM; a warp takes kGroup consecutive reads at a time and
// lane j of chunk c owns the k-mer that starts at base 32c+j of the current read.
//   STAGE   the group's words go from global to a per-warp shared-memory strip with coalesced 4-byte
//           loads (as many reads of the group as fit the strip at once; a read that is longer than
//           the strip is walked in windows).
//   KEYS    two LDS + funnel shifts give the 64 bits that start at the lane's base: the top 32 are
//           its 16-mer (minimizer candidate), the top 60 its forward k-mer (:481-517); __brev gives
//           the reverse complement, min() the canonical key (:528).  A block covers 4 chunks
//           (128 k-mers, a whole 150-base read): the 16-mer hashes of 5 chunk positions are computed
//           once (the fifth is the 14-lane halo of the fourth).
//   MINIM   sliding minimum of the 16-mer hashes over 15 positions: 4 shuffle rounds (1,2,4,7).
//   LOOKUP  sector = group(M) | sector(key); all four chunks (128 k-mers) request the 16-byte first
//           half of their sector (the high words of its three keys) before the first is consumed;
//           only a lane whose high word matches reads the second half, from L1 (kid_table2.cuh).
//           Lanes that share a minimizer share a 128-byte line, so a warp-wide load touches ~5 lines
//           instead of 32.
//   FOLD    hits are rare; a ballot finds them and the warp folds them strictly in position order
//           with kid_msca (the fold is order dependent).
//   COUNT   seen bit (atomicOr on the per-sample bitmap) for hits with taxon > 1 (:596-603),
//           gcount[final]++ (:613) in a shared-memory histogram flushed once per block.
#include "kid_kernels.cuh"

#include <cstdlib>
#include <type_traits>

namespace {

constexpr int kWarpsPerBlock = KID_CLASSIFY_THREADS / 32;
constexpr int kStripWords = 128;  // code words a staged run of reads may span (2048 bases)
constexpr int kStripPad = 12;     // zero words behind them: halo lanes never need a bounds check
constexpr int kMaskWords = 72;    // "a run of 30 valid bases starts here" bits of ONE read / window
constexpr int kGroupDefault = 8;  // consecutive reads a warp fetches the metadata of together (6: 16.90 ms, 8: 16.78, 12: 16.76)
constexpr int kLongStarts = 1920; // k-mer starts per window of a read longer than the strip (120 words)

struct WarpStrip {
    uint32_t codes[kStripWords + kStripPad];
    uint32_t kmask[kMaskWords];
};

__device__ __forceinline__ uint32_t ld_words(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// kmask of one read (or one window of a long read) from its validity words v[0..nv): bit 31-(j&31) of
// word j>>5 says that bases j..j+29 are all ACGT (cpos == KSIZE at base j+29, :526)
__device__ __forceinline__ void build_kmask(WarpStrip &strip, const uint32_t *v, int nv, int n_out, int lane)
{
    __syncwarp();
    for (int i = lane; i < n_out; i += 32) {
        const uint32_t a = i < nv ? __ldg(v + i) : 0u, b = i + 1 < nv ? __ldg(v + i + 1) : 0u;
        uint64_t x = ((uint64_t)a << 32) | b;
        x &= x << 1; x &= x << 2; x &= x << 4; x &= x << 8; // runs of 16
        x &= x << 14;                                       // runs of 30
        strip.kmask[i] = (uint32_t)(x >> 32);
    }
    __syncwarp();
}

struct ScanState { // all warp-uniform
    uint32_t fin;       // running final taxon of the read (:588-595)
    unsigned n_lookups; // getHash calls (:529); one warp's share of a launch stays far below 2^32
    unsigned n_hits;
};

// LOOKUP .. FOLD for N chunks of 32 k-mers whose sector indices are known.  All N first halves (high
// words of the three keys, 4 registers) are in flight before the first is consumed; a lane reads the
// second half of its sector - from L1 - only if a high word matched (kid_table2.cuh).
template <int N, bool kMerged, bool kNoL1>
__device__ __forceinline__ void lookup_chunks(const KidPackedParams &p, const Kid2TableView &tab, const uint64_t *key,
                                              const uint32_t *sec, const bool *act, ScanState &st)
{
    const unsigned full = 0xFFFFFFFFu;
    uint4 h[N];
    // Inactive lanes (k-mer with an N, or outside the read) read sector 0 instead of branching; their
    // result is ignored.
#pragma unroll
    for (int u = 0; u < N; u++) {
        const uint32_t si = act[u] ? sec[u] : 0u;
        if (kNoL1) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                                : "=r"(h[u].x), "=r"(h[u].y), "=r"(h[u].z), "=r"(h[u].w) : "l"(tab.sectors + 2 * (uint64_t)si));
        else h[u] = kid2_load_half(tab.sectors + 2 * (uint64_t)si);
    }
    uint32_t taxon[N]; // 0 = miss (taxon 0 is never stored)
    uint32_t again = 0, any_cand = 0; // bit u of again: sector full without the key - it may live further on
#pragma unroll
    for (int u = 0; u < N; u++) {
        const uint32_t khi = (uint32_t)(key[u] >> 32) | 0x80000000u;
        const uint32_t cand = act[u] ? kid2_candidates(h[u], khi) : 0u;
        any_cand |= cand;
        taxon[u] = 0;
        again |= act[u] && kid2_full(h[u]) ? (1u << u) : 0u;
        if (!kMerged && __any_sync(full, cand != 0)) { // hits are rare: most chunks skip this
            if (cand) {
                int j = 0;
                uint32_t tx = 0;
                if (kid2_verify(tab.sectors + 2 * (uint64_t)sec[u], h[u], cand, (uint32_t)key[u], tx, j)) {
                    taxon[u] = tx;
                    again &= ~(1u << u);
                    if (tx > 1) { // :596-603 - fire and forget, the OR is idempotent
                        const uint64_t slot = KID2_SLOTS_PER_SECTOR * (uint64_t)sec[u] + (uint64_t)j;
                        atomicOr(p.seen + (slot >> 5), 1u << (slot & 31));
                    }
                }
            }
        }
    }
    if (kMerged && __any_sync(full, any_cand != 0)) { // hits are rare: reads without any skip this
#pragma unroll
        for (int u = 0; u < N; u++) {
            const uint32_t cand = act[u] ? kid2_candidates(h[u], (uint32_t)(key[u] >> 32) | 0x80000000u) : 0u;
            if (cand) {
                int j = 0;
                uint32_t tx = 0;
                if (kid2_verify(tab.sectors + 2 * (uint64_t)sec[u], h[u], cand, (uint32_t)key[u], tx, j)) {
                    taxon[u] = tx;
                    again &= ~(1u << u);
                    if (tx > 1) {
                        const uint64_t slot = KID2_SLOTS_PER_SECTOR * (uint64_t)sec[u] + (uint64_t)j;
                        atomicOr(p.seen + (slot >> 5), 1u << (slot & 31));
                    }
                }
            }
        }
    }
    if (__any_sync(full, again != 0)) { // the few lanes that met a full sector: their next sectors, all loads first
#pragma unroll
        for (int u = 0; u < N; u++) kid2_load_half_if(tab.sectors + 2 * ((uint64_t)sec[u] + 1), h[u], (again >> u) & 1u);
#pragma unroll
        for (int u = 0; u < N; u++) {
            if ((again >> u) & 1u) {
                const uint32_t khi = (uint32_t)(key[u] >> 32) | 0x80000000u;
                const uint32_t cand = kid2_candidates(h[u], khi);
                int j = 0;
                uint32_t tx = 0;
                uint64_t slot = 0;
                bool found = false;
                if (cand && kid2_verify(tab.sectors + 2 * ((uint64_t)sec[u] + 1), h[u], cand, (uint32_t)key[u], tx, j)) {
                    found = true;
                    slot = KID2_SLOTS_PER_SECTOR * ((uint64_t)sec[u] + 1) + (uint64_t)j; // slack sectors follow the last home sector
                } else if (kid2_full(h[u])) { // rare: third sector and on
                    tx = kid2_lookup_from(tab, sec[u], key[u], 2, slot);
                    found = tx != 0;
                }
                if (found) {
                    taxon[u] = tx;
                    if (tx > 1) atomicOr(p.seen + (slot >> 5), 1u << (slot & 31));
                }
            }
        }
    }
    // FOLD, strictly in position order
    if (kMerged) {
        uint32_t any_hit = 0;
#pragma unroll
        for (int u = 0; u < N; u++) any_hit |= taxon[u];
        if (!__any_sync(full, any_hit != 0)) return;
    }
#pragma unroll
    for (int u = 0; u < N; u++) {
        unsigned m = __ballot_sync(full, taxon[u] != 0);
        if (m) {
            st.n_hits += __popc(m);
            do { // ordered left fold over the hits of this chunk (:588-595)
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t tj = __shfl_sync(full, taxon[u], src);
                if (st.fin > 0) { if (tj != st.fin) st.fin = kid_msca(p.tree, tj, st.fin); }
                else st.fin = tj;
            } while (m);
        }
    }
}

// KEYS .. FOLD over the k-mers that start at positions [c, c+128) of a read whose first base sits at
// staged index tbase; `last` = its last k-mer start; kmask bits apply only when `flagged`
template <int kInFlight, int kMM, int kVar>
__device__ __forceinline__ void scan_block(const KidPackedParams &p, const Kid2TableView &tab, const WarpStrip &strip,
                                           int tbase, int c, int last, bool flagged, int lane, ScanState &st)
{
    const unsigned full = 0xFFFFFFFFu;
    const int t0 = tbase + c + lane; // staged index of this lane's base in chunk 0; chunk u: + 32u
    const uint32_t *cw = strip.codes + (t0 >> 4);
    const int sh = (t0 & 15) * 2;
    constexpr int kW = kMM == 16 ? 10 : 11;
    uint32_t W[kW];
#pragma unroll
    for (int i = 0; i < kW; i++) W[i] = cw[i];
    // minimizer candidates -> sliding minima: the 32-bit hash of a 16-mer, or the (order, identity) pair of a 20-mer
    using MinT = typename std::conditional<kMM == 16, uint32_t, uint64_t>::type;
    MinT cm[5];
    uint64_t key[4];
#pragma unroll
    for (int u = 0; u < 5; u++) {
        const uint32_t hi = __funnelshift_l(W[2 * u + 1], W[2 * u], sh); // bases 0..15 at this position
        const uint32_t rc = kid_rc16(hi);
        uint32_t lo = 0, rl = 0;
        if (u < 4 || kMM != 16) {
            lo = __funnelshift_l(W[2 * u + 2], W[2 * u + 1], sh); // bases 16..31
            rl = kid_rc16(lo);
        }
        if (kMM == 16) {
            cm[u] = (MinT)kid_mm_hash_canon(min(hi, rc));
        } else {
            // canonical 20-mer as (top 32 bits, low 8 bits): forward = bases 0..19, reverse complement =
            // that of bases 16..19 in front of that of bases 0..15 (kid_table2.cuh)
            const uint32_t ft = hi, fl = lo >> 24;
            const uint32_t rt = ((rl & 0xFFu) << 24) | (rc >> 8), rlo = rc & 0xFFu;
            const bool fwd = ft < rt || (ft == rt && fl < rlo);
            cm[u] = (MinT)kid_mm20_pair(fwd ? ft : rt, fwd ? fl : rlo);
        }
        if (u < 4) {
            // forward key = first 30 of the 32 bases at this position; the reverse complement of 32
            // bases is rc16(low half) : rc16(high half), its low 60 bits that of the first 30 bases
            const uint64_t kf = (((uint64_t)hi << 32) | lo) >> 4;
            const uint64_t kr = (((uint64_t)rl << 32) | rc) & KID_MASK60;
            key[u] = kf < kr ? kf : kr; // :528
        }
    }
    // MINIM: window minimum over the 15 (m = 16) or 11 (m = 20) candidate positions of a 30-mer by
    // doubling: 1 + 2 + 4 + 7 resp. 1 + 2 + 4 + 3
#pragma unroll
    for (int step = 0; step < 4; step++) {
        const int d = step == 0 ? 1 : step == 1 ? 2 : step == 2 ? 4 : (kMM == 16 ? 7 : 3);
        const int src = lane + d; // shfl takes the source lane modulo 32
        const bool wrap = lane + d >= 32;
        MinT s[5];
#pragma unroll
        for (int u = 0; u < 5; u++) s[u] = __shfl_sync(full, cm[u], src);
#pragma unroll
        for (int u = 0; u < 4; u++) cm[u] = min(cm[u], wrap ? s[u + 1] : s[u]);
        if (step < 3) cm[4] = min(cm[4], wrap ? (MinT)~(MinT)0 : s[4]);
    }
    // which of the 128 positions are k-mer starts: inside the read (the first nvb positions of the
    // block) and, for a read with non-ACGT bases, the start of a run of 30 valid bases (kmask: position
    // 32u+j in bit 31-j of word u, warp-uniform)
    bool act[4];
    uint32_t sec[4];
    const int nvb = last - c + 1; // >= 1
    if ((kVar & 2) && !flagged) {
#pragma unroll
        for (int u = 0; u < 4; u++) act[u] = lane < nvb - 32 * u;
        st.n_lookups += (unsigned)min(nvb, 128); // each is one getHash call (:529)
    } else {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int nv = nvb - 32 * u;
            uint32_t m = nv >= 32 ? 0xFFFFFFFFu : (nv <= 0 ? 0u : 0xFFFFFFFFu << (32 - nv));
            if (flagged) m &= strip.kmask[(c >> 5) + u];
            st.n_lookups += __popc(m);
            act[u] = (int32_t)(m << lane) < 0;
        }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const uint32_t grp = ((uint32_t)cm[u] * 0x9E3779B1u) >> tab.line_shift; // (m = 20: the winner's identity word)
        sec[u] = (grp << tab.sub_bits) | (kid_key_hash32(key[u]) >> (32 - tab.sub_bits));
    }
    // every sector index is known before the first load goes out, so that the compiler cannot slide
    // the tail of MINIM between the loads
    asm volatile("" ::"r"(sec[0]), "r"(sec[1]), "r"(sec[2]), "r"(sec[3]));
    if (kInFlight == 4) {
        lookup_chunks<4, (kVar & 1) != 0, (kVar & 4) != 0>(p, tab, key, sec, act, st);
    } else {
        lookup_chunks<2, (kVar & 1) != 0, (kVar & 4) != 0>(p, tab, key, sec, act, st);
        if (c + 64 <= last) lookup_chunks<2, (kVar & 1) != 0, (kVar & 4) != 0>(p, tab, key + 2, sec + 2, act + 2, st); // warp-uniform
    }
}

template <bool SMEM_HIST, int kInFlight, int kMM, int kGroup, int kVar>
__global__ void __launch_bounds__(KID_CLASSIFY_THREADS, 1)
kid_classify3_kernel(const KidPackedParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpStrip *strips = reinterpret_cast<WarpStrip *>(smem_raw);
    int *hist = reinterpret_cast<int *>(smem_raw + sizeof(WarpStrip) * kWarpsPerBlock);

    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    WarpStrip &strip = strips[warp_in_block];
    const Kid2TableView tab = p.table2;

    if (SMEM_HIST) {
        for (int i = threadIdx.x; i < p.tree.n_taxa; i += blockDim.x) hist[i] = 0;
        __syncthreads();
    }
    if (lane < kStripPad) strip.codes[kStripWords + lane] = 0;

    ScanState st;
    st.fin = 0;
    st.n_lookups = 0;
    st.n_hits = 0;

    auto finish_read = [&](size_t r, bool kept) { // per-read output and gcount[final]++ (:613)
        if (lane == 0) {
            if (p.out_taxon) p.out_taxon[r] = kept ? (int32_t)st.fin : -1;
            if (kept) {
                if (SMEM_HIST) atomicAdd(&hist[st.fin], 1);
                else atomicAdd(&p.gcount[st.fin], 1);
            }
        }
    };

    const size_t n_groups = (p.n_reads + kGroup - 1) / kGroup;
    const size_t warps_total = (size_t)gridDim.x * kWarpsPerBlock;
    for (size_t grp = (size_t)blockIdx.x * kWarpsPerBlock + warp_in_block; grp < n_groups; grp += warps_total) {
        const size_t r0 = grp * kGroup;
        const int nr = (int)min((size_t)kGroup, p.n_reads - r0);
        // lane i holds read r0+i: first word (relative to the batch), length, flag; lane nr the end
        uint32_t w_first = 0, tlen = 0;
        if (lane <= nr) {
            const uint2 m = __ldg(p.meta + r0 + lane);
            w_first = m.x;
            tlen = m.y;
        }
        const bool my_flag = (w_first & KID_PK_FLAG) != 0;
        w_first = (w_first & ~KID_PK_FLAG) - p.word_bias;
        // words this read occupies: codes + validity words if flagged
        const uint32_t my_words = lane < nr ? ((tlen + 15) >> 4) + (my_flag ? (tlen + 31) >> 5 : 0u) : 0u;
        const uint32_t w_end = w_first + my_words;

        int s = 0;
        while (s < nr) {
            const uint32_t base = __shfl_sync(full, w_first, s);
            // reads s..e-1 are staged together: the largest run whose words end inside the strip
            const unsigned fits = __ballot_sync(full, lane >= s && lane < nr && w_end - base <= (uint32_t)kStripWords);
            const int e = s + __popc(fits); // word indices do not decrease, so the lanes that fit are s..e-1
            if (e > s) {
                const uint32_t span = __shfl_sync(full, w_end, e - 1) - base;
                __syncwarp();
                for (uint32_t i = lane; i < span; i += 32) strip.codes[i] = ld_words(p.words + base + i);
                __syncwarp();
                for (int i = s; i < e; i++) {
                    const int tl = (int)__shfl_sync(full, tlen, i);
                    const uint32_t wf = __shfl_sync(full, w_first, i);
                    const bool flagged = __shfl_sync(full, (int)my_flag, i) != 0;
                    const bool kept = tl > KID_KSIZE; // stop - start >= KSIZE (:755)
                    st.fin = 0;
                    if (kept) {
                        const int last = tl - KID_KSIZE;
                        if (flagged) build_kmask(strip, p.words + wf + ((tl + 15) >> 4), (tl + 31) >> 5, (last >> 5) + 1, lane);
                        const int tbase = (int)(wf - base) * 16;
                        for (int c = 0; c <= last; c += 128) scan_block<kInFlight, kMM, kVar>(p, tab, strip, tbase, c, last, flagged, lane, st);
                    }
                    finish_read(r0 + i, kept);
                }
                s = e;
                continue;
            }
            // ---- read s alone does not fit the strip: windows of kLongStarts k-mer starts
            {
                const int tl = (int)__shfl_sync(full, tlen, s);
                const uint32_t wf = __shfl_sync(full, w_first, s);
                const bool flagged = __shfl_sync(full, (int)my_flag, s) != 0;
                const bool kept = tl > KID_KSIZE;
                st.fin = 0;
                if (kept) {
                    const int last = tl - KID_KSIZE;
                    const int cwords = (tl + 15) >> 4, vwords = (tl + 31) >> 5;
                    for (int wb = 0; wb <= last; wb += kLongStarts) {
                        const int wlast = min(kLongStarts - 1, last - wb);
                        const int w0 = wb >> 4, nw = min(kStripWords, cwords - w0);
                        __syncwarp();
                        for (int i = lane; i < kStripWords; i += 32) strip.codes[i] = i < nw ? ld_words(p.words + wf + w0 + i) : 0u;
                        __syncwarp();
                        if (flagged)
                            build_kmask(strip, p.words + wf + cwords + (wb >> 5), vwords - (wb >> 5), (wlast >> 5) + 1, lane);
                        for (int c = 0; c <= wlast; c += 128) scan_block<kInFlight, kMM, kVar>(p, tab, strip, 0, c, wlast, flagged, lane, st);
                    }
                }
                finish_read(r0 + s, kept);
                s++;
            }
        }
    }

    if (lane == 0) {
        if (st.n_lookups) atomicAdd(p.counters + 0, st.n_lookups);
        if (st.n_hits) atomicAdd(p.counters + 1, st.n_hits);
    }
    if (SMEM_HIST) {
        __syncthreads();
        for (int i = threadIdx.x; i < p.tree.n_taxa; i += blockDim.x) {
            const int cnt = hist[i];
            if (cnt) atomicAdd(&p.gcount[i], cnt);
        }
    }
}

template <bool H, int F, int MM, int G, int V>
cudaError_t launch_one(const KidPackedParams &p, int sm_count, cudaStream_t stream)
{
    const size_t smem = sizeof(WarpStrip) * kWarpsPerBlock + (H ? (size_t)p.tree.n_taxa * 4 : 0);
    auto kern = kid_classify3_kernel<H, F, MM, G, V>;
    cudaError_t err = cudaSuccess;
    if (smem > 48 * 1024) {
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
    }
    // shared memory and L1 are one 256 KB array: reserve only what the block uses, the rest caches
    // the taxonomy rows
    if (getenv("KID_NO_CARVEOUT") == nullptr) {
        int pct = (int)(((smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
        if (pct > 100) pct = 100;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct); // a hint: failure is harmless
    }
    size_t blocks = (size_t)sm_count; // persistent: one block per SM
    const size_t need = ((p.n_reads + G - 1) / G + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks > need) blocks = need;
    if (blocks == 0) return cudaSuccess;
    kern<<<(unsigned)blocks, KID_CLASSIFY_THREADS, smem, stream>>>(p);
    KID_COUNT_LAUNCH();
    return cudaGetLastError();
}

template <int F, int MM, int G, int V>
cudaError_t launch_hist(const KidPackedParams &p, int sm_count, cudaStream_t stream)
{
    const bool hist = (size_t)p.tree.n_taxa * 4 <= KID_SMEM_HIST_MAX_BYTES;
    return hist ? launch_one<true, F, MM, G, V>(p, sm_count, stream) : launch_one<false, F, MM, G, V>(p, sm_count, stream);
}

// measured (tools/gpu_runs/gpu_i.sh, 20 M reads): per-chunk votes + range masks 15.69 ms; one vote per block 16.27;
// lane compare instead of range masks 15.53; both 15.94; first halves bypassing L1 (bit 2): 16.22
constexpr int kVarDefault = 2;
template <int F, int G = kGroupDefault, int V = kVarDefault>
cudaError_t launch_variant(const KidPackedParams &p, int sm_count, cudaStream_t stream)
{
    return p.table2.mm == KID_MM20 ? launch_hist<F, KID_MM20, G, V>(p, sm_count, stream) : launch_hist<F, 16, G, V>(p, sm_count, stream);
}

} // namespace

cudaError_t kid_launch_classify3(const KidPackedParams &p, int sm_count, cudaStream_t stream)
{
#ifdef KID_TUNE_VARIANTS // experiment builds only: KID_TUNE=<n>
    static const int tune = getenv("KID_TUNE") ? atoi(getenv("KID_TUNE")) : 0;
    switch (tune) {
    case 1: return launch_variant<2>(p, sm_count, stream);     // 2 chunks in flight, then the other 2
    case 2: return launch_variant<4, kGroupDefault, 0>(p, sm_count, stream); // range masks instead of the lane compare
    case 3: return launch_variant<4, kGroupDefault, 3>(p, sm_count, stream); // one vote for all candidates / hits of a block
    case 4: return launch_variant<4, kGroupDefault, 6>(p, sm_count, stream); // first halves bypass L1
    default: break;
    }
#endif
    return launch_variant<4>(p, sm_count, stream); // all 4 chunks of a block in flight
}
