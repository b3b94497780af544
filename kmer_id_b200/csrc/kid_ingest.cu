// kid_ingest.cu - gzip FASTQ files read ON THE DEVICE: the compressed bytes cross PCIe, everything that
// process_fqgz (newkmer_10nx.cpp:762-816) and process_qual (:714-760) do per line happens in kernels.
//
//   inflate   kid_inflate.cuh: block finder (a warp per piece), speculative inflate to 16-bit symbols
//             (a thread per piece), the chain check on the host, window maps in groups, marker resolve,
//             CRC-32 of every gzip member in 4 KiB chunks;
//   frame     newline positions (count, prefix sum, write), per line: strip one '\r', empty lines do
//             not advance the 4-line state (:786-802) -> the state of a line is the number of non-empty
//             lines before it mod 4 (the rule of host/read_reader.cpp), i.e. a second prefix sum;
//             record r = non-empty lines 4r (header), 4r+1 (bases), 4r+3 (qualities);
//   gather    bases and qualities of all records into the text batch kid_pack_kernel consumes;
//   classify  kid_pack_kernel + kid_classify3_kernel over the whole file, per-read taxa back to the host.
//
// Anything this path does not reproduce byte for byte - a file zlib has to judge (no gzip header,
// trailing bytes, a bad check value, fixed/stored-only streams), a line of >= 16 KiB (fatal in the
// reference), a quality line shorter than its read (the reference aborts), more text than device memory
// holds (the whole file is inflated at once) - makes kid_fastq_load_gz_file return KID_EUNSUPPORTED BEFORE anything was
// counted; the caller then reads the file with the host reader, whose error behaviour is the reference's.
#include "kid_internal.cuh"
#include "kid_inflate_chain.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <new>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

using namespace kidz;

#define KID_TRY(call)                \
    do {                             \
        const int rc_ = (call);      \
        if (rc_ != KID_OK) return rc_; \
    } while (0)

namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kFindWarps = 8;
constexpr int kFindLanes = 8; // lanes of a warp that run the full header parse at a time (each needs its own tables)
constexpr int kFindWarpEntries = kFindLanes * kFindTabEntries + 256; // + two queues of 64 candidate offsets
constexpr int kInflateWarps = 4;
constexpr uint32_t kRefLineLimit = 0x4000;                  // BUFLEN, newkmer_10nx.cpp:85
constexpr uint32_t kErrLongLine = 1u, kErrShortQual = 2u;

// ---------------------------------------------------------------------------------------- find
// One warp per piece k >= 1.  Three tests of rising cost, each run on 32 survivors of the one before at
// a time so that the lanes stay busy: the header's first 13 bits (all lanes, consecutive bit positions;
// ~22 % pass), a complete code-length code (~1 in 200), the full header parse (kid_inflate.cuh).
// Queues hold bit offsets relative to the piece, in increasing order; the first position that passes
// the full test wins.
__device__ __forceinline__ uint32_t queue_push(uint32_t *q, uint32_t qn, bool pass, uint32_t value, uint32_t lane)
{
    const unsigned m = __ballot_sync(kFull, pass);
    if (pass) q[qn + __popc(m & ((1u << lane) - 1u))] = value;
    __syncwarp();
    return qn + __popc(m);
}
__device__ __forceinline__ uint32_t queue_drop32(uint32_t *q, uint32_t qn, uint32_t lane) // qn <= 64
{
    const uint32_t nb = min(qn, 32u);
    const uint32_t keep = lane + 32u < qn ? q[lane + 32u] : 0u;
    __syncwarp();
    if (lane + 32u < qn) q[lane] = keep;
    __syncwarp();
    return qn - nb;
}

// Stages two and three over the queues; `more` = the piece has positions left (then the queues only have
// to get below 32 entries), else everything is drained.  Returns the piece-relative bit of the first
// position that passes the full test, or kNoBit.  Kept out of line: inlined, its header parser pushed the
// scan loop's own variables into local memory (90 instructions per 32 positions instead of 25).
constexpr uint32_t kNoBit = 0xFFFFFFFFu;
struct FindQueues {
    uint32_t found, n1, n2;
};
__device__ __noinline__ FindQueues find_drain(const uint32_t *w, const uint32_t *wp, uint64_t from, uint32_t *q1, uint32_t *q2, uint16_t *tabmem,
                                              uint32_t n1, uint32_t n2, bool more, bool text_only, uint32_t lane)
{
    Tab<kFindLanes> tab{ tabmem + (lane % kFindLanes) };
    uint32_t found = kNoBit;
    for (;;) {
        while ((n1 >= 32 || (!more && n1 > 0)) && n2 < 32) {
            const uint32_t c = lane < n1 ? q1[lane] : 0u;
            const bool pass = lane < n1 && block_start_cl_complete(wp, c, peek32(wp, c));
            n2 = queue_push(q2, n2, pass, c, lane); // <= 63
            n1 = queue_drop32(q1, n1, lane);
        }
        while (n2 >= 32 || (!more && n1 == 0 && n2 > 0)) {
            const uint32_t nb = min(n2, 32u);
            for (uint32_t c0 = 0; c0 < nb && found == kNoBit; c0 += kFindLanes) { // rare: kFindLanes candidates at a time
                const uint32_t i = c0 + lane;
                const uint32_t c = lane < kFindLanes && i < nb ? q2[i] : 0u;
                const bool ok = lane < kFindLanes && i < nb && is_block_start(w, from + c, tab, text_only);
                const unsigned m = __ballot_sync(kFull, ok);
                if (m) found = __shfl_sync(kFull, c, __ffs(m) - 1);
                __syncwarp();
            }
            if (found != kNoBit) break;
            n2 = queue_drop32(q2, n2, lane);
        }
        if (found != kNoBit || !(n1 >= 32 || (!more && n1 > 0))) break;
    }
    return FindQueues{ found, n1, n2 };
}

__global__ void __launch_bounds__(kFindWarps * 32)
kidz_find_kernel(const uint32_t *w, uint64_t size, uint64_t piece_bytes, uint32_t n_pieces, bool text_only, uint64_t *start_bit)
{
    extern __shared__ uint16_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t k = 1u + blockIdx.x * kFindWarps + warp;
    if (k >= n_pieces) return;
    uint16_t *wmem = smem + warp * kFindWarpEntries;
    uint32_t *q1 = reinterpret_cast<uint32_t *>(wmem + kFindLanes * kFindTabEntries), *q2 = q1 + 64;
    const uint64_t from = (uint64_t)k * piece_bytes * 8; // a multiple of 32: piece_bytes is one of 4
    const uint32_t *wp = w + (from >> 5);
    const uint32_t n_bits = (uint32_t)(min(size, ((uint64_t)k + 1) * piece_bytes) * 8 - from);
    const uint32_t n_iter = (n_bits + 31u) >> 5;
    uint32_t n1 = 0, n2 = 0, found = kNoBit;
    uint32_t mine = 0, after = 0; // 32 words of the piece, one per lane, and the word behind them
    for (uint32_t it = 0; it < n_iter; it++) { // positions 32 it .. 32 it + 31, one per lane
        if ((it & 31u) == 0) {
            mine = wp[it + lane];
            after = wp[it + 32u];
        }
        const uint32_t lo = __shfl_sync(kFull, mine, (int)(it & 31u));
        const uint32_t nx = __shfl_sync(kFull, mine, (int)((it + 1u) & 31u));
        const uint32_t v = __funnelshift_r(lo, (it & 31u) == 31u ? after : nx, lane);
        const uint32_t r = it * 32u + lane;
        n1 = queue_push(q1, n1, r < n_bits && block_start_bits_plausible(v), r, lane);
        if (n1 >= 32) {
            const FindQueues q = find_drain(w, wp, from, q1, q2, wmem, n1, n2, true, text_only, lane);
            n1 = q.n1;
            n2 = q.n2;
            found = q.found;
            if (found != kNoBit) break;
        }
    }
    if (found == kNoBit) found = find_drain(w, wp, from, q1, q2, wmem, n1, n2, false, text_only, lane).found;
    if (lane == 0) start_bit[k] = found == kNoBit ? ~0ull : from + found;
}

// ------------------------------------------------------------------------------------- inflate
struct InflateArgs {
    const uint32_t *w;
    uint64_t size, piece_bytes, first_block_bit;
    uint32_t n_pieces;
    const uint64_t *start_bit;
    uint16_t *syms;
    uint32_t slot;
    PieceResult *res;
    uint32_t *counter;
};

// ---- one decoder per WARP ---------------------------------------------------------------------------
// Huffman decoding is serial and a file only has so many deflate blocks to start from (one per ~25 KiB of
// compressed FASTQ: some 10^4 per sample), fewer than the GPU holds warps, so what counts is the time of ONE
// decoder.  Measured on the way here (3844 pieces of 32 KiB, 634 MB of text):
//   32 decoders per warp, free running      290 ms  the lanes drift apart and execute one at a time
//   32 decoders per warp in lockstep         41 ms  (Inflater::step + __syncwarp) every path of a step is
//                                                   executed, 9 of 32 lanes active (r2_inflate_lockstep_ncu.txt)
//   one decoder on one lane of a warp        28 ms  118 warp instructions per symbol: issue bound
//   one decoder on all lanes of a warp (this kernel)
// Inside a block every lane decodes the token (a literal, or a whole length/distance pair) that would
// start at "its" bit: lane i looks at bit bp + i.  The tokens that really are there form a chain from
// lane 0 (token at i ends where the next one starts) which is walked with shuffles; the lanes on the
// chain write their output at positions given by a prefix sum, matches as 32-wide fills.  ~10 symbols per
// round instead of one.  Block headers, tokens with codes longer than the fast tables, the gzip trailer
// and everything unusual go through the plain decoder (Inflater) on lane 0.
enum WarpRound : int { kRoundEob = 0, kRoundBad = 1, kRoundOverflow = 2, kRoundSlowToken = 3 };

__device__ __forceinline__ int warp_symbols(const uint32_t *w, uint64_t size_bits, Tab<1> t, uint16_t *out, uint64_t &bp, int32_t &pos,
                                            int32_t floor, int32_t cap, uint32_t lane)
{
    for (;;) {
        if (bp > size_bits) return kRoundBad; // ran off the end of the file
        // 64 bits from bit bp + lane on
        const uint64_t b = bp + lane;
        const uint32_t *p = w + (b >> 5);
        const uint32_t sh = (uint32_t)b & 31u;
        const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
        const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
        const uint64_t v = (uint64_t)lo | ((uint64_t)hi << 32);
        // the token that starts here, if one does
        const uint32_t e = t.at(kOffLitFast + (int)(lo & ((1u << kLitRoot) - 1u)));
        uint32_t bits = e & 15u, outlen = 1, code = e >> 4;
        uint32_t kind = 0; // 0 literal, 1 match, 2 end of block, 3 not for this path
        if (bits == 0 || code > 285u) kind = 3;
        else if (code == 256u) { kind = 2; outlen = 0; }
        else if (code > 256u) {
            uint32_t len;
            if (code < 265u) len = code - 254u;
            else if (code == 285u) len = 258u;
            else {
                const uint32_t x = code - 261u, eb = x >> 2;
                len = 3u + ((4u + (x & 3u)) << eb) + ((uint32_t)(v >> bits) & ((1u << eb) - 1u));
                bits += eb;
            }
            const uint32_t de = t.at(kOffDistFast + (int)((uint32_t)(v >> bits) & ((1u << kDistRoot) - 1u)));
            const uint32_t ds = de >> 4;
            if ((de & 15u) == 0 || ds > 29u) kind = 3;
            else {
                bits += de & 15u;
                uint32_t dist;
                if (ds < 4u) dist = ds + 1u;
                else {
                    const uint32_t eb = (ds >> 1) - 1u;
                    dist = 1u + ((2u + (ds & 1u)) << eb) + ((uint32_t)(v >> bits) & ((1u << eb) - 1u));
                    bits += eb;
                }
                kind = 1;
                outlen = len;
                code = kCopyFlag | (dist - 1u);
            }
        }
        // the chain of real tokens from lane 0: one shuffle per token brings where the next one starts
        // (relative to bp, <= 31 + 48), its kind and how many positions it writes
        const uint32_t link = (lane + bits) | (kind << 7) | (outlen << 9);
        uint32_t cur = 0, total = 0;
        int32_t at = 0;
        bool on = false, eob = false;
        while (cur < 32u) {
            const uint32_t l = __shfl_sync(kFull, link, (int)cur);
            const uint32_t kd = (l >> 7) & 3u;
            if (kd == 3u) break;
            if (lane == cur) {
                on = true;
                at = pos + (int32_t)total;
            }
            total += l >> 9;
            cur = l & 127u;
            if (kd == 2u) {
                eob = true;
                break;
            }
        }
        if (total == 0 && !eob && cur == 0) return kRoundSlowToken; // the token at bp itself needs the plain decoder
        if (total > (uint32_t)(cap - pos)) return kRoundOverflow;
        const bool match = on && kind == 1u;
        // a match may not reach before the member / before any history
        if (__any_sync(kFull, match && (int32_t)((code & 0x7fffu) + 1u) > at - floor)) return kRoundBad;
        if (on && kind == 0u) out[at] = (uint16_t)code;
        // every position of a match gets "copy from dist back": a short match by its own lane, a long one
        // by the whole warp, 32 positions per store
        if (match && outlen <= 8u)
            for (uint32_t i = 0; i < outlen; i++) out[at + (int32_t)i] = (uint16_t)code;
        unsigned mm = __ballot_sync(kFull, match && outlen > 8u);
        while (mm) {
            const int src = __ffs(mm) - 1;
            mm &= mm - 1;
            const int32_t mat = __shfl_sync(kFull, at, src);
            const uint32_t mlen = __shfl_sync(kFull, outlen, src);
            const uint32_t mcode = __shfl_sync(kFull, code, src);
            for (uint32_t i = lane; i < mlen; i += 32) out[mat + (int32_t)i] = (uint16_t)mcode;
        }
        pos += (int32_t)total;
        bp += cur;
        if (eob) return kRoundEob;
    }
}

// the whole piece: lane 0 owns the plain decoder and lends its state to the warp inside blocks
__device__ __forceinline__ void warp_inflate_piece(const uint32_t *w, uint64_t size, uint64_t start_bit, uint64_t stop_bit, int floor0,
                                                   uint16_t *out, uint32_t out_cap, Tab<1> tab, PieceResult *res, uint32_t lane)
{
    Inflater<1> d;
    d.state = Inflater<1>::kDone;
    d.pos = 0;
    d.floor = 0;
    d.in.next = 0;
    d.in.cnt = 0;
    if (lane == 0) d.start(w, size, start_bit, stop_bit, floor0, out, out_cap, tab, res);
    for (;;) {
        if (lane == 0 && d.state == Inflater<1>::kAtBoundary) d.block_header();
        __syncwarp(); // the tables lane 0 built are in shared memory
        if (__shfl_sync(kFull, d.state, 0) == Inflater<1>::kDone) break;
        if (__shfl_sync(kFull, d.state, 0) == Inflater<1>::kAtBoundary) continue; // e.g. a stored block: the next header
        uint64_t bp = __shfl_sync(kFull, (unsigned long long)d.in.pos(), 0);
        int32_t pos = __shfl_sync(kFull, d.pos, 0);
        const int32_t floor = __shfl_sync(kFull, d.floor, 0);
        const int rc = warp_symbols(w, size * 8, tab, out, bp, pos, floor, (int32_t)out_cap, lane);
        __syncwarp();
        if (lane == 0) {
            d.pos = pos;
            d.in.seek(w, bp);
            if (rc == kRoundEob) d.end_of_block();
            else if (rc == kRoundBad) d.refuse(kPieceBadData);
            else if (rc == kRoundOverflow) d.refuse(kPieceOverflow);
            else {
                d.symbol();
                while (d.fill_left) d.write_fill();
            }
        }
    }
}

__global__ void __launch_bounds__(kInflateWarps * 32)
kidz_inflate_kernel(const InflateArgs a)
{
    extern __shared__ uint16_t smem[];
    const uint32_t lane = threadIdx.x & 31u;
    Tab<1> tab{ smem + (threadIdx.x >> 5) * kTabEntries };
    for (;;) {
        uint32_t k = 0;
        if (lane == 0) k = atomicAdd(a.counter, 1u);
        k = __shfl_sync(kFull, k, 0);
        if (k >= a.n_pieces) break;
        PieceResult *r = a.res + k;
        const uint64_t start = k == 0 ? a.first_block_bit : a.start_bit[k];
        if (start == ~0ull) {
            if (lane == 0) {
                r->start_bit = r->end_bit = 0;
                r->n_out = r->n_ends = r->eof = 0;
                r->status = kPieceNoStart;
            }
            continue;
        }
        const uint64_t stop = min(a.size, ((uint64_t)k + 1) * a.piece_bytes) * 8;
        warp_inflate_piece(a.w, a.size, start, stop, k == 0 ? 0 : kWin, a.syms + (size_t)k * a.slot, a.slot, tab, r, lane);
    }
}

// one piece again, from the bit its predecessor stopped at (the chain walk asks for it)
__global__ void __launch_bounds__(32)
kidz_inflate_one_kernel(const InflateArgs a, uint32_t k, uint64_t start)
{
    extern __shared__ uint16_t smem[];
    Tab<1> tab{ smem };
    const uint64_t stop = min(a.size, ((uint64_t)k + 1) * a.piece_bytes) * 8;
    warp_inflate_piece(a.w, a.size, start, stop, kWin, a.syms + (size_t)k * a.slot, a.slot, tab, a.res + k, threadIdx.x & 31u);
}

// ------------------------------------------------------------------------------------------ copy
// resolve_copies (kid_inflate.cuh) with one warp per piece, two chunks of 32 positions per round: a copy
// whose source lies before the round's first position is one gather load (everything before it is final),
// one whose source lies inside its own chunk is chased through the lanes with shuffles (pointer jumping,
// <= 5 rounds), one of the second chunk whose source lies in the first takes it from there by shuffle.
//
// chase of the copies whose source lies inside the same chunk of 32 positions: sl = source lane (below
// the lane's own), done = the lane holds its final symbol
__device__ __forceinline__ uint32_t chase_in_chunk(uint32_t v, uint32_t sl, bool done)
{
    unsigned pending = __ballot_sync(kFull, !done);
    while (pending) {
        const uint32_t vv = __shfl_sync(kFull, v, (int)sl);
        const int dd = __shfl_sync(kFull, (int)done, (int)sl);
        const uint32_t ss = __shfl_sync(kFull, sl, (int)sl);
        if (!done) {
            if (dd) {
                v = vv;
                done = true;
            } else sl = ss;
        }
        pending = __ballot_sync(kFull, !done);
    }
    return v;
}

__global__ void __launch_bounds__(64)
kidz_copy_kernel(uint16_t *syms, uint32_t slot, const PieceResult *res, uint32_t k0, uint32_t n_pieces)
{
    const uint32_t k = k0 + blockIdx.x * 2u + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31u;
    if (k >= n_pieces || res[k].status != kPieceOk) return;
    const uint32_t n = res[k].n_out;
    uint16_t *out = syms + (size_t)k * slot;
    // two chunks (A: p.., B: p + 32..) per round, so that two gathers per lane are in flight; the codes of the
    // next round are loaded before this round's gathers come back
    uint32_t a_next = lane < n ? out[lane] : 0u, b_next = 32u + lane < n ? out[32u + lane] : 0u;
    for (uint32_t p = 0; p < n; p += 64) {
        const uint32_t xa = a_next, xb = b_next;
        if (p + 64u + lane < n) a_next = out[p + 64u + lane];
        if (p + 96u + lane < n) b_next = out[p + 96u + lane];
        const bool have_a = p + lane < n, have_b = p + 32u + lane < n;
        uint32_t va = xa, sla = lane, vb = xb, slb = lane;
        bool done_a = true, done_b = true, from_a = false;
        if (have_a && (xa & kCopyFlag)) {
            const int32_t src = (int32_t)(p + lane) - (int32_t)((xa & 0x7fffu) + 1u);
            if (src < 0) va = (uint32_t)(256 + kWin + src);
            else if ((uint32_t)src < p) va = __ldcg(out + src);
            else {
                done_a = false;
                sla = (uint32_t)src - p;
            }
        }
        if (have_b && (xb & kCopyFlag)) {
            const int32_t src = (int32_t)(p + 32u + lane) - (int32_t)((xb & 0x7fffu) + 1u);
            if (src < 0) vb = (uint32_t)(256 + kWin + src);
            else if ((uint32_t)src < p) vb = __ldcg(out + src);
            else if ((uint32_t)src < p + 32u) { // in chunk A: after A is final
                from_a = true;
                slb = (uint32_t)src - p;
            } else {
                done_b = false;
                slb = (uint32_t)src - p - 32u;
            }
        }
        va = chase_in_chunk(va, sla, done_a);
        const uint32_t fa = __shfl_sync(kFull, va, (int)(from_a ? slb : lane));
        if (from_a) vb = fa;
        vb = chase_in_chunk(vb, slb, done_b);
        if (have_a) out[p + lane] = (uint16_t)va;
        if (have_b) out[p + 32u + lane] = (uint16_t)vb;
        __syncwarp(); // the next round's gathers may read these
    }
}

// --------------------------------------------------------------------------------- window maps
// Accepted pieces in groups of G.  Inside a group, serially: pm[j] = the 32 KiB before piece j as a
// function of the 32 KiB before the group (a byte, or 256 + index into that window); gm[g] = the same for
// the window after the group's last piece.
__global__ void __launch_bounds__(1024)
kidz_maps_kernel(const uint16_t *syms, uint32_t slot, const uint32_t *pieces, const PieceResult *res, uint32_t n_acc, uint32_t group,
                 uint16_t *pm, uint16_t *gm)
{
    extern __shared__ uint16_t smem[];
    uint16_t *cur = smem, *nxt = smem + kWin;
    const uint32_t g = blockIdx.x;
    for (int i = threadIdx.x; i < kWin; i += 1024) cur[i] = (uint16_t)(256 + i);
    __syncthreads();
    const uint32_t j1 = min(n_acc, (g + 1) * group);
    for (uint32_t j = g * group; j < j1; j++) {
        uint16_t *pmj = pm + (size_t)j * kWin;
        const uint32_t k = pieces[j];
        const uint32_t n = res[k].n_out;
        const uint16_t *s = syms + (size_t)k * slot;
        for (uint32_t i = threadIdx.x; i < (uint32_t)kWin; i += 1024) {
            pmj[i] = cur[i];
            uint16_t v;
            if (n < (uint32_t)kWin && i < (uint32_t)kWin - n) v = cur[i + n];
            else {
                const uint16_t x = s[n - (uint32_t)kWin + i];
                v = x < 256 ? x : cur[x - 256];
            }
            nxt[i] = v;
        }
        __syncthreads();
        uint16_t *t = cur;
        cur = nxt;
        nxt = t;
    }
    uint16_t *gmg = gm + (size_t)g * kWin;
    for (int i = threadIdx.x; i < kWin; i += 1024) gmg[i] = cur[i];
}

// gw[g] = the 32 KiB before group g as bytes; one block walks the groups
__global__ void __launch_bounds__(1024)
kidz_group_windows_kernel(const uint16_t *gm, uint32_t n_groups, uint8_t *gw)
{
    extern __shared__ uint16_t smem[];
    uint8_t *cur = reinterpret_cast<uint8_t *>(smem), *nxt = cur + kWin;
    for (int i = threadIdx.x; i < kWin; i += 1024) cur[i] = 0; // before the file: never referenced by a valid stream
    __syncthreads();
    for (uint32_t g = 0; g < n_groups; g++) {
        uint8_t *gwg = gw + (size_t)g * kWin;
        const uint16_t *gmg = gm + (size_t)g * kWin;
        for (int i = threadIdx.x; i < kWin; i += 1024) {
            gwg[i] = cur[i];
            const uint16_t x = gmg[i];
            nxt[i] = x < 256 ? (uint8_t)x : cur[x - 256];
        }
        __syncthreads();
        uint8_t *t = cur;
        cur = nxt;
        nxt = t;
    }
}

// symbols -> bytes: a marker goes through the piece's map, then through its group's window
__global__ void __launch_bounds__(512)
kidz_resolve_kernel(const uint16_t *syms, uint32_t slot, const uint32_t *pieces, const PieceResult *res, const uint64_t *text_off,
                    const uint16_t *pm, const uint8_t *gw, uint32_t group, uint8_t *text)
{
    const uint32_t j = blockIdx.x;
    const uint32_t k = pieces[j];
    const uint32_t n = res[k].n_out;
    const uint16_t *s = syms + (size_t)k * slot;
    const uint16_t *pmj = pm + (size_t)j * kWin;
    const uint8_t *win = gw + (size_t)(j / group) * kWin;
    uint8_t *o = text + text_off[j];
    for (uint32_t i = threadIdx.x; i < n; i += 512) {
        uint32_t v = s[i];
        if (v >= 256) {
            v = pmj[v - 256];
            if (v >= 256) v = win[v - 256];
        }
        o[i] = (uint8_t)v;
    }
}

// ------------------------------------------------------------------------------------- CRC-32
// consts: 256 table entries, then the 32 powers x^(2^k)
__global__ void __launch_bounds__(256)
kidz_crc_kernel(const uint8_t *text, uint64_t n_text, const Member *members, uint32_t n_members, const uint32_t *consts, uint32_t *acc)
{
    __shared__ uint32_t tab[256];
    __shared__ uint32_t pow2[32];
    tab[threadIdx.x] = consts[threadIdx.x];
    if (threadIdx.x < 32) pow2[threadIdx.x] = consts[256 + threadIdx.x];
    __syncthreads();
    const uint64_t c0 = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * 4096;
    if (c0 >= n_text) return;
    const uint64_t cend = min(n_text, c0 + 4096);
    uint32_t lo = 0, hi = n_members; // first member that ends after c0
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (members[mid].end <= c0) lo = mid + 1; else hi = mid;
    }
    uint32_t m = lo;
    uint64_t a = c0;
    while (a < cend && m < n_members) {
        const uint64_t mend = members[m].end;
        if (mend <= a) { m++; continue; } // an empty member
        const uint64_t b = min(cend, mend);
        uint32_t crc = 0xffffffffu;
        uint64_t i = a;
        if ((a & 15u) == 0) // the usual case, a whole chunk inside one member: 16 bytes per load
            for (; i + 16 <= b; i += 16) {
                const uint4 q = *reinterpret_cast<const uint4 *>(text + i);
                const uint32_t wv[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    uint32_t x = wv[k];
#pragma unroll
                    for (int j = 0; j < 4; j++, x >>= 8) crc = tab[(crc ^ x) & 0xffu] ^ (crc >> 8);
                }
            }
        for (; i < b; i++) crc = tab[(crc ^ text[i]) & 0xffu] ^ (crc >> 8);
        crc ^= 0xffffffffu;
        atomicXor(acc + m, crc_mulmod(crc_x8n(mend - b, pow2), crc));
        a = b;
    }
}

// ----------------------------------------------------------------------------- prefix sums (u32)
// exclusive prefix sums of n = min(*n_dev, n_bound) items, 4096 per block: block sums, one block over
// the block sums, then every block again.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_sums, uint32_t &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += t;
    }
    __syncthreads(); // warp_sums may still be read from a previous call
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t before = 0, all = 0;
    for (int i = 0; i < n_warps; i++) {
        const uint32_t t = warp_sums[i];
        if (i < warp) before += t;
        all += t;
    }
    total = all;
    return before + incl - v;
}

// the same over 64-bit values (the spine of the scans: sums of block sums pass 2^32 with more than 4 G bases)
__device__ __forceinline__ uint64_t block_exclusive_scan64(uint64_t v, uint64_t *warp_sums, uint64_t &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    uint64_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t t = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += t;
    }
    __syncthreads();
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint64_t before = 0, all = 0;
    for (int i = 0; i < n_warps; i++) {
        const uint64_t t = warp_sums[i];
        if (i < warp) before += t;
        all += t;
    }
    total = all;
    return before + incl - v;
}

__global__ void __launch_bounds__(1024)
kidz_scan_sums_kernel(const uint32_t *in, const uint64_t *n_dev, uint32_t n_bound, uint64_t *block_sums)
{
    __shared__ uint32_t ws[32];
    const uint32_t n = n_dev ? (uint32_t)min(*n_dev, (uint64_t)n_bound) : n_bound;
    const uint64_t i0 = ((uint64_t)blockIdx.x * 1024 + threadIdx.x) * 4;
    uint32_t v = 0;
    for (int q = 0; q < 4; q++)
        if (i0 + q < n) v += in[i0 + q];
    uint32_t total;
    block_exclusive_scan(v, ws, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
kidz_scan_spine_kernel(uint64_t *block_sums, uint32_t n_blocks, uint64_t *total_out)
{
    __shared__ uint64_t ws[32];
    uint64_t carry = 0;
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += 1024) {
        const uint32_t i = b0 + threadIdx.x;
        const uint64_t v = i < n_blocks ? block_sums[i] : 0u;
        uint64_t total;
        const uint64_t ex = block_exclusive_scan64(v, ws, total);
        if (i < n_blocks) block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <class OutT>
__global__ void __launch_bounds__(1024)
kidz_scan_apply_kernel(const uint32_t *in, const uint64_t *n_dev, uint32_t n_bound, const uint64_t *block_sums, OutT *out)
{
    __shared__ uint32_t ws[32];
    const uint32_t n = n_dev ? (uint32_t)min(*n_dev, (uint64_t)n_bound) : n_bound;
    const uint64_t i0 = ((uint64_t)blockIdx.x * 1024 + threadIdx.x) * 4;
    uint32_t x[4], v = 0;
    for (int q = 0; q < 4; q++) {
        x[q] = i0 + q < n ? in[i0 + q] : 0u;
        v += x[q];
    }
    uint32_t total;
    uint64_t run = block_sums[blockIdx.x] + block_exclusive_scan(v, ws, total);
    for (int q = 0; q < 4; q++) {
        if (i0 + q < n) out[i0 + q] = (OutT)run;
        run += x[q];
        if (i0 + q + 1 == n) out[n] = (OutT)run; // the end offset
    }
    if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) out[0] = 0;
}

// ------------------------------------------------------------------------------------ framing
// newline flags of the 16 bytes a thread owns (bit i = byte i), bytes at or beyond n_text masked off
__device__ __forceinline__ uint32_t newline_mask16(const uint8_t *text, uint64_t n_text, uint64_t at)
{
    if (at >= n_text) return 0;
    const uint4 q = *reinterpret_cast<const uint4 *>(text + at);
    const uint32_t wv[4] = { q.x, q.y, q.z, q.w };
    uint32_t mask = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t x = wv[k] ^ 0x0A0A0A0Au;
        const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu); // 0x80 in every byte that is '\n'
        mask |= (((z >> 7) & 1u) | ((z >> 14) & 2u) | ((z >> 21) & 4u) | ((z >> 28) & 8u)) << (4 * k);
    }
    const uint64_t left = n_text - at;
    if (left < 16) mask &= (1u << left) - 1u;
    return mask;
}

__global__ void __launch_bounds__(256)
kidz_nl_count_kernel(const uint8_t *text, uint64_t n_text, uint32_t *tile_count)
{
    __shared__ uint32_t ws[32];
    const uint64_t at = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * 16;
    uint32_t total;
    block_exclusive_scan((uint32_t)__popc(newline_mask16(text, n_text, at)), ws, total);
    if (threadIdx.x == 0) tile_count[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256)
kidz_nl_write_kernel(const uint8_t *text, uint64_t n_text, const uint32_t *tile_base, uint64_t *nlpos)
{
    __shared__ uint32_t ws[32];
    const uint64_t at = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * 16;
    uint32_t mask = newline_mask16(text, n_text, at);
    uint32_t total;
    uint32_t idx = tile_base[blockIdx.x] + block_exclusive_scan((uint32_t)__popc(mask), ws, total);
    while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        nlpos[idx++] = at + (uint64_t)b;
    }
}

// line j = text[start, nlpos[j]) with start = nlpos[j-1] + 1; one trailing '\r' does not count (:786-787)
__device__ __forceinline__ void line_span(const uint8_t *text, const uint64_t *nlpos, uint32_t j, uint64_t &start, uint32_t &len,
                                          unsigned long long *err)
{
    start = j ? nlpos[j - 1] + 1u : 0u;
    const uint64_t e = nlpos[j];
    const uint64_t l = e - start;
    if (l >= kRefLineLimit) atomicOr(err, (unsigned long long)kErrLongLine); // "Buffer to small for input line lengths" (:773)
    len = (uint32_t)min(l, (uint64_t)kRefLineLimit);
    if (len > 0 && text[e - 1] == '\r') len--;
}

struct RecordArrays { // one entry per record; positions in the text
    uint64_t *hdr_start, *seq_start, *qual_start;
    uint32_t *hdr_len, *seq_len, *qual_len;
};

__global__ void __launch_bounds__(256)
kidz_line_count_kernel(const uint8_t *text, uint64_t n_text, const uint64_t *nlpos, uint32_t n_lines, uint32_t *block_count,
                       unsigned long long *err)
{
    __shared__ uint32_t ws[32];
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    uint32_t nonempty = 0;
    if (j < n_lines) {
        uint64_t start;
        uint32_t len;
        line_span(text, nlpos, j, start, len, err);
        nonempty = len > 0;
        // the unterminated tail of the stream is dropped (:812-813) unless it overflows the line buffer
        if (j == n_lines - 1 && n_text - (nlpos[j] + 1) >= kRefLineLimit) atomicOr(err, (unsigned long long)kErrLongLine);
    }
    uint32_t total;
    block_exclusive_scan(nonempty, ws, total);
    if (threadIdx.x == 0) block_count[blockIdx.x] = total;
}

// misc[0] = non-empty lines in total (written by the spine scan); records = that / 4 (:788-802)
__global__ void __launch_bounds__(256)
kidz_line_write_kernel(const uint8_t *text, const uint64_t *nlpos, uint32_t n_lines, const uint32_t *block_base, const uint64_t *misc,
                       RecordArrays rec, unsigned long long *err)
{
    __shared__ uint32_t ws[32];
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    uint32_t nonempty = 0, len = 0;
    uint64_t start = 0;
    if (j < n_lines) {
        line_span(text, nlpos, j, start, len, err);
        nonempty = len > 0;
    }
    uint32_t total;
    const uint32_t idx = block_base[blockIdx.x] + block_exclusive_scan(nonempty, ws, total);
    if (!nonempty) return;
    const uint32_t r = idx >> 2, n_records = (uint32_t)(misc[0] >> 2);
    if (r >= n_records) return; // an incomplete last record is never handed to process_qual
    switch (idx & 3u) {
    case 0: rec.hdr_start[r] = start; rec.hdr_len[r] = len; break;
    case 1: rec.seq_start[r] = start; rec.seq_len[r] = len; break;
    case 3: rec.qual_start[r] = start; rec.qual_len[r] = len; break;
    default: break;
    }
}

__global__ void kidz_set_records_kernel(uint64_t *misc)
{
    misc[1] = misc[0] >> 2;
}

// a warp per record: its bases and the qualities under them, appended to the text batch
__global__ void __launch_bounds__(256)
kidz_gather_kernel(const uint8_t *text, RecordArrays rec, const uint64_t *misc, const uint64_t *off, uint8_t *seq, uint8_t *qual,
                   unsigned long long *err)
{
    const uint32_t n = (uint32_t)misc[1];
    const int lane = threadIdx.x & 31;
    for (uint32_t r = blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += gridDim.x * 8) {
        const uint32_t len = rec.seq_len[r];
        if (rec.qual_len[r] < len) { // qual.at(stop) throws (:729): the reference aborts
            if (lane == 0) atomicOr(err, (unsigned long long)kErrShortQual);
            continue;
        }
        const uint8_t *s = text + rec.seq_start[r], *q = text + rec.qual_start[r];
        const uint64_t o = off[r];
        for (uint32_t i = lane; i < len; i += 32) {
            seq[o + i] = s[i];
            qual[o + i] = q[i];
        }
    }
}

// ----------------------------------------------------------------------------- _reads.txt fetch
// lens[2i] = header length, lens[2i+1] = trimmed bases of read idx[i]; pos[i] = where they go
__global__ void __launch_bounds__(1024)
kidz_fetch_plan_kernel(const uint32_t *idx, uint32_t n, RecordArrays rec, const uint32_t *span, uint32_t *lens, uint32_t *pos)
{
    __shared__ uint32_t ws[32];
    uint32_t carry = 0;
    for (uint32_t b0 = 0; b0 < n; b0 += 1024) {
        const uint32_t i = b0 + threadIdx.x;
        uint32_t v = 0;
        if (i < n) {
            const uint32_t r = idx[i];
            const uint32_t hl = rec.hdr_len[r], bl = span[2 * r + 1] - span[2 * r] + 1u;
            lens[2 * i] = hl;
            lens[2 * i + 1] = bl;
            v = hl + bl;
        }
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, ws, total);
        if (i < n) pos[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) pos[n] = carry;
}

__global__ void __launch_bounds__(256)
kidz_fetch_copy_kernel(const uint8_t *text, const uint32_t *idx, uint32_t n, RecordArrays rec, const uint32_t *span, const uint32_t *lens,
                       const uint32_t *pos, uint8_t *out)
{
    const int lane = threadIdx.x & 31;
    for (uint32_t i = blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += gridDim.x * 8) {
        const uint32_t r = idx[i];
        const uint32_t hl = lens[2 * i], bl = lens[2 * i + 1];
        uint8_t *o = out + pos[i];
        const uint8_t *h = text + rec.hdr_start[r], *b = text + rec.seq_start[r] + span[2 * r];
        for (uint32_t c = lane; c < hl; c += 32) o[c] = h[c];
        for (uint32_t c = lane; c < bl; c += 32) o[hl + c] = b[c];
    }
}

// ---------------------------------------------------------------------------------- host side
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 16 + 256;
        const cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        const cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocPortable);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

size_t env_size(const char *name, size_t def)
{
    const char *e = getenv(name);
    return e && *e ? (size_t)strtoull(e, nullptr, 10) : def;
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

} // namespace

enum { kPhRead = 0, kPhFind, kPhInflate, kPhChain, kPhResolve, kPhFrame, kPhClassify, kPhFetch, kPhCount };

// A compressed file on its way to the device: read in 8 MiB chunks into two page-locked staging buffers,
// each copied asynchronously while the next is being read.  A kid_fastq has two, so that the next file can
// be fetched (kid_fastq_prefetch_gz_file, a helper thread) while the kernels work on the current one.
struct GzLoader {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = { nullptr, nullptr };
    PinBuf stage[2];
    DevBuf gz;
    uint64_t size = 0, first_block_bit = 0;
    bool text_only = false; // the file's first block codes no byte below 9 or from 128 on: see is_block_start
    const char *why = "";   // with KID_EUNSUPPORTED

    int open_streams()
    {
        if (!stream) KID_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        for (cudaEvent_t &e : ev)
            if (!e) KID_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        return KID_OK;
    }
    int read(const char *path)
    {
        why = "";
        KID_TRY(open_streams());
        const int fd = open(path, O_RDONLY);
        if (fd < 0) { why = "cannot open"; return KID_EUNSUPPORTED; }
        struct FdCloser { int fd; ~FdCloser() { close(fd); } } closer{ fd };
        struct stat sb;
        if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) { why = "not a regular file"; return KID_EUNSUPPORTED; }
        size = (uint64_t)sb.st_size;
        if (size < 18) { why = "shorter than a gzip member"; return KID_EUNSUPPORTED; }
        if (size >= ((uint64_t)1 << 36)) { why = "larger than 64 GiB"; return KID_EUNSUPPORTED; }
        const size_t gz_cap = (size_t)((size + 3) & ~3ull) + 1024; // the block finder loads whole 128-byte rows
        KID_CUDA(gz.reserve(gz_cap));
        const size_t tail0 = (size_t)size & ~(size_t)3; // zero padding behind the file: the bit readers run a few words past it
        KID_CUDA(cudaMemsetAsync(gz.as<uint8_t>() + tail0, 0, gz_cap - tail0, stream));
        const size_t chunk = (size_t)8 << 20;
        for (int i = 0; i < 2; i++) KID_CUDA(stage[i].reserve(chunk));
        uint64_t done = 0;
        int which = 0;
        bool used[2] = { false, false };
        const bool timing = getenv("KID_GZ_GPU_TIMING") != nullptr;
        double t_wait = 0, t_read = 0, t_copy = 0, t_mark = timing ? now_s() : 0;
        auto lap = [&](double &acc) {
            if (!timing) return;
            const double t = now_s();
            acc += t - t_mark;
            t_mark = t;
        };
        while (done < size) {
            if (used[which]) KID_CUDA(cudaEventSynchronize(ev[which]));
            lap(t_wait);
            uint8_t *dst = stage[which].as<uint8_t>();
            size_t got = 0;
            const size_t want = (size_t)std::min<uint64_t>(chunk, size - done);
            while (got < want) {
                const ssize_t r = pread(fd, dst + got, want - got, (off_t)(done + got));
                if (r <= 0) { why = "read error"; return KID_EUNSUPPORTED; }
                got += (size_t)r;
            }
            lap(t_read);
            if (done == 0) {
                const uint64_t hl = gzip_header_len(dst, std::min<uint64_t>(size, want), 0);
                if (!hl) { why = "no gzip header"; return KID_EUNSUPPORTED; }
                first_block_bit = hl * 8;
                // the block finder's text-only filter is used if the file's own first block passes it
                text_only = false;
                if (want >= hl + 1024 && (peek32(reinterpret_cast<const uint32_t *>(dst), first_block_bit) & 6u) == 4u) {
                    uint16_t tabmem[kFindTabEntries];
                    text_only = is_block_start(reinterpret_cast<const uint32_t *>(dst), first_block_bit, Tab<1>{ tabmem }, true);
                }
            }
            KID_CUDA(cudaMemcpyAsync(gz.as<uint8_t>() + done, dst, want, cudaMemcpyHostToDevice, stream));
            KID_CUDA(cudaEventRecord(ev[which], stream));
            used[which] = true;
            which ^= 1;
            done += want;
            lap(t_copy);
        }
        KID_CUDA(cudaStreamSynchronize(stream));
        lap(t_wait);
        if (timing) fprintf(stderr, "[kid_fastq] %s: pread %.3f s, copy calls %.3f s, waiting for copies %.3f s\n", path, t_read, t_copy, t_wait);
        return KID_OK;
    }
    void release()
    {
        gz.release();
        for (PinBuf &b : stage) b.release();
        for (cudaEvent_t e : ev)
            if (e) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }
};

struct kid_fastq {
    const kid_db *db = nullptr;
    cudaStream_t stream = nullptr;
    GzLoader loader[2]; // [cur]: the file the kernels read; the other one: the file being prefetched
    int cur = 0;
    std::thread pre_thread;
    std::string pre_path;
    bool pre_active = false;
    int pre_rc = KID_OK;
    bool attrs_set = false;
    size_t piece_bytes = 32768, expand = 8, group = 64;
    DevBuf start, res, syms, counter, pieces, text_off, members, acc, consts, pm, gm, gw, text, tiles, tilecnt, linecnt, nlpos, recs, off, seq, qual,
        misc, words, meta, taxon, span, fidx, flens, fpos, fdata;
    PinBuf h_res, h_small, h_fetch_lens, h_fetch_data;
    bool consts_ready = false;
    // the file that is loaded
    size_t n_reads = 0;
    uint64_t n_text = 0, n_bases = 0, gz_bytes = 0;
    uint32_t n_lines = 0;
    bool loaded = false;
    RecordArrays rec{};
    size_t n_pieces = 0, n_redo = 0, n_covered = 0, n_members = 0;
    double phase_s[kPhCount] = { 0 };
    std::vector<PieceResult> res_host;
};

namespace {

int set_kernel_attrs(kid_fastq *f)
{
    if (f->attrs_set) return KID_OK;
    KID_CUDA(cudaFuncSetAttribute(kidz_find_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFindWarps * kFindWarpEntries * 2));
    KID_CUDA(cudaFuncSetAttribute(kidz_inflate_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 75));
    KID_CUDA(cudaFuncSetAttribute(kidz_maps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kWin * 2));
    KID_CUDA(cudaFuncSetAttribute(kidz_group_windows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kWin));
    f->attrs_set = true;
    return KID_OK;
}

// exclusive prefix sums of in[0..n) -> out[0..n] on the stream; the total also goes to *total_out (device)
template <class OutT>
int scan_u32(kid_fastq *f, const uint32_t *in, const uint64_t *n_dev, uint32_t n_bound, OutT *out, uint64_t *total_out)
{
    const uint32_t nb = (n_bound + 4095) / 4096 + 1; // one block more than needed keeps n_bound == 0 simple
    KID_CUDA(f->tiles.reserve(sizeof(uint64_t) * ((size_t)nb + 1)));
    uint64_t *bs = f->tiles.as<uint64_t>();
    kidz_scan_sums_kernel<<<nb, 1024, 0, f->stream>>>(in, n_dev, n_bound, bs);
    kidz_scan_spine_kernel<<<1, 1024, 0, f->stream>>>(bs, nb, total_out);
    kidz_scan_apply_kernel<OutT><<<nb, 1024, 0, f->stream>>>(in, n_dev, n_bound, bs, out);
    for (int i = 0; i < 3; i++) KID_COUNT_LAUNCH();
    KID_CUDA(cudaGetLastError());
    return KID_OK;
}

int unsupported(const char *path, const char *why)
{
    return kid_fail(KID_EUNSUPPORTED, "%s: left to the host reader (%s)", path, why);
}

} // namespace

extern "C" {

int kid_fastq_create(const kid_db *db, kid_fastq **out)
{
    if (!db || !out) return kid_fail(KID_EINVAL, "kid_fastq_create: NULL argument");
    *out = nullptr;
    DeviceGuard guard(db->device);
    kid_fastq *f = new (std::nothrow) kid_fastq;
    if (!f) return kid_fail(KID_ENOMEM, "kid_fastq_create: host allocation failed");
    f->db = db;
    f->piece_bytes = std::max<size_t>(4096, env_size("KID_GZ_GPU_PIECE", 32768) & ~(size_t)3);
    f->expand = std::max<size_t>(2, env_size("KID_GZ_GPU_EXPAND", 8));
    f->group = std::max<size_t>(1, env_size("KID_GZ_GPU_GROUP", 64));
    const cudaError_t e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        kid_fastq_free(f);
        return kid_fail(KID_ECUDA, "kid_fastq_create: %s", cudaGetErrorString(e));
    }
    *out = f;
    return KID_OK;
}

void kid_fastq_free(kid_fastq *f)
{
    if (!f) return;
    DeviceGuard guard(f->db->device);
    if (f->pre_active) f->pre_thread.join();
    if (f->stream) cudaStreamSynchronize(f->stream);
    for (GzLoader &l : f->loader) l.release();
    for (DevBuf *b : { &f->start, &f->res, &f->syms, &f->counter, &f->pieces, &f->text_off, &f->members, &f->acc, &f->consts,
                       &f->pm, &f->gm, &f->gw, &f->text, &f->tiles, &f->tilecnt, &f->linecnt, &f->nlpos, &f->recs, &f->off, &f->seq, &f->qual, &f->misc, &f->words,
                       &f->meta, &f->taxon, &f->span, &f->fidx, &f->flens, &f->fpos, &f->fdata })
        b->release();
    for (PinBuf *b : { &f->h_res, &f->h_small, &f->h_fetch_lens, &f->h_fetch_data }) b->release();
    if (f->stream) cudaStreamDestroy(f->stream);
    delete f;
}

int kid_fastq_prefetch_gz_file(kid_fastq *f, const char *path)
{
    if (!f || !path) return kid_fail(KID_EINVAL, "kid_fastq_prefetch_gz_file: NULL argument");
    if (f->pre_active) {
        f->pre_thread.join();
        f->pre_active = false;
    }
    f->pre_path = path;
    f->pre_active = true;
    GzLoader *ld = &f->loader[1 - f->cur];
    const int device = f->db->device;
    const double t_call = now_s();
    f->pre_thread = std::thread([f, ld, device, t_call] {
        cudaSetDevice(device);
        const double t_go = now_s();
        f->pre_rc = ld->read(f->pre_path.c_str());
        if (getenv("KID_GZ_GPU_TIMING"))
            fprintf(stderr, "[kid_fastq] %s: read ahead, thread up after %.3f s, done after %.3f s\n", f->pre_path.c_str(), t_go - t_call, now_s() - t_call);
    });
    return KID_OK;
}

static int load_gz_file(kid_fastq *f, const char *path, size_t *n_reads);

int kid_fastq_load_gz_file(kid_fastq *f, const char *path, size_t *n_reads)
{
    if (!f || !path) return kid_fail(KID_EINVAL, "kid_fastq_load_gz_file: NULL argument");
    const int rc = load_gz_file(f, path, n_reads);
    if (rc != KID_ENOMEM) return rc;
    // an allocation failed after all (another tenant of the GPU, a file that expands more than estimated): the
    // big buffers go back and the host reader takes the file
    DeviceGuard guard(f->db->device);
    cudaStreamSynchronize(f->stream);
    cudaGetLastError();
    for (DevBuf *b : { &f->syms, &f->pm, &f->gm, &f->gw, &f->text, &f->nlpos, &f->recs, &f->off, &f->seq, &f->qual, &f->words, &f->meta,
                       &f->taxon, &f->span })
        b->release();
    f->loaded = false;
    return unsupported(path, "a device allocation failed");
}

static int load_gz_file(kid_fastq *f, const char *path, size_t *n_reads)
{
    if (n_reads) *n_reads = 0;
    f->loaded = false;
    f->n_reads = 0;
    DeviceGuard guard(f->db->device);
    KID_TRY(set_kernel_attrs(f));
    cudaStream_t st = f->stream;
    for (double &p : f->phase_s) p = 0;
    double t0 = now_s();

    // ---- the compressed file -> device (already there if it was prefetched)
    GzLoader *ld = nullptr;
    if (f->pre_active) {
        f->pre_thread.join();
        if (getenv("KID_GZ_GPU_TIMING")) fprintf(stderr, "[kid_fastq] %s: waited %.3f s for the read-ahead\n", path, now_s() - t0);
        f->pre_active = false;
        if (f->pre_path == path) {
            ld = &f->loader[1 - f->cur];
            if (f->pre_rc == KID_EUNSUPPORTED) return unsupported(path, ld->why);
            if (f->pre_rc != KID_OK) return kid_fail(f->pre_rc, "%s: reading it ahead failed", path);
            f->cur = 1 - f->cur;
        }
    }
    if (!ld) {
        ld = &f->loader[f->cur];
        const int rc = ld->read(path);
        if (rc == KID_EUNSUPPORTED) return unsupported(path, ld->why);
        if (rc != KID_OK) return rc;
    }
    const uint64_t size = ld->size, first_block_bit = ld->first_block_bit;
    const uint64_t P = f->piece_bytes;
    const size_t n_pieces = (size_t)((size + P - 1) / P);
    const uint64_t slot64 = P * f->expand + (192u << 10);
    if (slot64 >= 0x7ff00000ull || n_pieces >= 0x7fffffffull) return unsupported(path, "piece size out of range");
    const uint32_t slot = (uint32_t)slot64;
    {
        size_t free_b = 0, total_b = 0;
        const double tq = now_s();
        KID_CUDA(cudaMemGetInfo(&free_b, &total_b));
        if (getenv("KID_GZ_GPU_TIMING")) fprintf(stderr, "[kid_fastq] %s: cudaMemGetInfo %.3f s\n", path, now_s() - tq);
        // symbols (2 bytes each, reserved per piece) and window maps for sure; text, line and record tables, the text
        // batch and its packed form (~3.5x the text) for FASTQ's usual 6x expansion - checked again once the text's
        // size is known
        const uint64_t need = (uint64_t)n_pieces * slot * 2 + (uint64_t)n_pieces * kWin * 2 + size * 21;
        const uint64_t have = free_b + f->syms.cap + f->pm.cap + f->text.cap + f->seq.cap + f->qual.cap + f->words.cap + f->nlpos.cap + f->recs.cap;
        if (need + ((uint64_t)2 << 30) > have) return unsupported(path, "not enough device memory for the whole file");
    }
    f->gz_bytes = size;
    f->n_pieces = n_pieces;
    double t1 = now_s();
    f->phase_s[kPhRead] = t1 - t0;

    // ---- find + inflate
    KID_CUDA(f->start.reserve(sizeof(uint64_t) * n_pieces));
    KID_CUDA(f->res.reserve(sizeof(PieceResult) * n_pieces));
    KID_CUDA(f->syms.reserve(sizeof(uint16_t) * n_pieces * (size_t)slot));
    KID_CUDA(f->counter.reserve(sizeof(uint32_t)));
    KID_CUDA(f->h_res.reserve(sizeof(PieceResult) * n_pieces));
    KID_CUDA(cudaMemsetAsync(f->counter.p, 0, sizeof(uint32_t), st));
    const uint32_t *w = ld->gz.as<uint32_t>();
    if (n_pieces > 1) {
        const unsigned blocks = (unsigned)((n_pieces - 1 + kFindWarps - 1) / kFindWarps);
        kidz_find_kernel<<<blocks, kFindWarps * 32, kFindWarps * kFindWarpEntries * 2, st>>>(w, size, P, (uint32_t)n_pieces, ld->text_only, f->start.as<uint64_t>());
        KID_COUNT_LAUNCH();
        KID_CUDA(cudaGetLastError());
    }
    if (getenv("KID_GZ_GPU_TIMING")) { KID_CUDA(cudaStreamSynchronize(st)); }
    double t2 = now_s();
    f->phase_s[kPhFind] = t2 - t1;
    InflateArgs ia;
    ia.w = w;
    ia.size = size;
    ia.piece_bytes = P;
    ia.first_block_bit = first_block_bit;
    ia.n_pieces = (uint32_t)n_pieces;
    ia.start_bit = f->start.as<uint64_t>();
    ia.syms = f->syms.as<uint16_t>();
    ia.slot = slot;
    ia.res = f->res.as<PieceResult>();
    ia.counter = f->counter.as<uint32_t>();
    {
        const unsigned blocks = (unsigned)std::min<size_t>((n_pieces + kInflateWarps - 1) / kInflateWarps, (size_t)f->db->sm_count * (64 / kInflateWarps));
        kidz_inflate_kernel<<<blocks, kInflateWarps * 32, kInflateWarps * kTabEntries * 2, st>>>(ia);
        KID_COUNT_LAUNCH();
        KID_CUDA(cudaGetLastError());
    }
    if (getenv("KID_GZ_GPU_TIMING")) {
        KID_CUDA(cudaStreamSynchronize(st));
        f->phase_s[kPhFetch] = -(now_s() - t2); // (diagnostic: the decode pass alone, shown negative in the fetch column)
    }
    kidz_copy_kernel<<<(unsigned)((n_pieces + 1) / 2), 64, 0, st>>>(f->syms.as<uint16_t>(), slot, f->res.as<PieceResult>(), 0u, (uint32_t)n_pieces);
    KID_COUNT_LAUNCH();
    KID_CUDA(cudaGetLastError());
    KID_CUDA(cudaMemcpyAsync(f->h_res.p, f->res.p, sizeof(PieceResult) * n_pieces, cudaMemcpyDeviceToHost, st));
    KID_CUDA(cudaStreamSynchronize(st));
    double t3 = now_s();
    f->phase_s[kPhInflate] = t3 - t2;

    // ---- chain
    f->res_host.assign(f->h_res.as<PieceResult>(), f->h_res.as<PieceResult>() + n_pieces);
    Chain chain;
    bool cuda_failed = false;
    const char *why = walk_chain(f->res_host, P, size, first_block_bit, [&](size_t k, uint64_t start) {
        kidz_inflate_one_kernel<<<1, 32, kTabEntries * 2, st>>>(ia, (uint32_t)k, start);
        kidz_copy_kernel<<<1, 64, 0, st>>>(f->syms.as<uint16_t>(), slot, f->res.as<PieceResult>(), (uint32_t)k, (uint32_t)k + 1u);
        KID_COUNT_LAUNCH();
        KID_COUNT_LAUNCH();
        cudaError_t e = cudaMemcpyAsync(f->h_res.as<PieceResult>() + k, f->res.as<PieceResult>() + k, sizeof(PieceResult), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { cuda_failed = true; return false; }
        f->res_host[k] = f->h_res.as<PieceResult>()[k];
        return true;
    }, env_size("KID_GZ_GPU_MAX_REDO", 8), chain);
    if (cuda_failed) return kid_fail(KID_ECUDA, "kid_fastq_load_gz_file: %s", cudaGetErrorString(cudaGetLastError()));
    if (why) return unsupported(path, why);
    const uint64_t T = chain.text_off.back();
    if (T >= ((uint64_t)1 << 36)) return unsupported(path, "more than 64 GiB of text");
    {   // what the rest needs on the device, now that the text's size is known: window maps, text, line and record
        // tables, the text batch, its packed form, per-read results
        size_t free_b = 0, total_b = 0;
        KID_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t need = (uint64_t)chain.pieces.size() * kWin * 2 + 3 * T + T / 2;
        const uint64_t have = free_b + f->pm.cap + f->text.cap + f->seq.cap + f->qual.cap + f->words.cap + f->nlpos.cap + f->recs.cap;
        if (need + ((uint64_t)1 << 30) > have) return unsupported(path, "not enough device memory for the whole file");
    }
    f->n_redo = chain.n_redo;
    f->n_covered = chain.n_covered;
    f->n_members = chain.members.size();
    const uint32_t M = (uint32_t)chain.pieces.size(), G = (uint32_t)f->group, NG = (M + G - 1) / G;
    double t4 = now_s();
    f->phase_s[kPhChain] = t4 - t3;

    // ---- maps, resolve, CRC, newline count
    KID_CUDA(f->pieces.reserve(sizeof(uint32_t) * M));
    KID_CUDA(f->text_off.reserve(sizeof(uint64_t) * ((size_t)M + 1)));
    KID_CUDA(f->members.reserve(sizeof(Member) * chain.members.size()));
    KID_CUDA(f->acc.reserve(sizeof(uint32_t) * chain.members.size()));
    KID_CUDA(f->pm.reserve((size_t)M * kWin * 2));
    KID_CUDA(f->gm.reserve((size_t)NG * kWin * 2));
    KID_CUDA(f->gw.reserve((size_t)NG * kWin));
    const size_t text_cap = (size_t)((T + 4095) & ~4095ull) + 4096;
    KID_CUDA(f->text.reserve(text_cap));
    KID_CUDA(f->misc.reserve(sizeof(uint64_t) * 16));
    KID_CUDA(f->h_small.reserve(sizeof(uint64_t) * 16 + sizeof(uint32_t) * chain.members.size()));
    if (!f->consts_ready) {
        uint32_t c[256 + 32];
        crc_make_table(c);
        crc_make_pow2(c + 256);
        KID_CUDA(f->consts.reserve(sizeof c));
        KID_CUDA(cudaMemcpyAsync(f->consts.p, c, sizeof c, cudaMemcpyHostToDevice, st));
        KID_CUDA(cudaStreamSynchronize(st)); // c lives on this stack frame
        f->consts_ready = true;
    }
    KID_CUDA(cudaMemcpyAsync(f->pieces.p, chain.pieces.data(), sizeof(uint32_t) * M, cudaMemcpyHostToDevice, st));
    KID_CUDA(cudaMemcpyAsync(f->text_off.p, chain.text_off.data(), sizeof(uint64_t) * ((size_t)M + 1), cudaMemcpyHostToDevice, st));
    KID_CUDA(cudaMemcpyAsync(f->members.p, chain.members.data(), sizeof(Member) * chain.members.size(), cudaMemcpyHostToDevice, st));
    KID_CUDA(cudaMemsetAsync(f->acc.p, 0, sizeof(uint32_t) * chain.members.size(), st));
    KID_CUDA(cudaMemsetAsync(f->misc.p, 0, sizeof(uint64_t) * 16, st));
    kidz_maps_kernel<<<NG, 1024, 2 * kWin * 2, st>>>(f->syms.as<uint16_t>(), slot, f->pieces.as<uint32_t>(), f->res.as<PieceResult>(), M, G,
                                                      f->pm.as<uint16_t>(), f->gm.as<uint16_t>());
    kidz_group_windows_kernel<<<1, 1024, 2 * kWin, st>>>(f->gm.as<uint16_t>(), NG, f->gw.as<uint8_t>());
    kidz_resolve_kernel<<<M, 512, 0, st>>>(f->syms.as<uint16_t>(), slot, f->pieces.as<uint32_t>(), f->res.as<PieceResult>(),
                                            f->text_off.as<uint64_t>(), f->pm.as<uint16_t>(), f->gw.as<uint8_t>(), G, f->text.as<uint8_t>());
    for (int i = 0; i < 3; i++) KID_COUNT_LAUNCH();
    KID_CUDA(cudaGetLastError());
    const uint32_t n_tiles = (uint32_t)((T + 4095) / 4096);
    if (T > 0) {
        const unsigned crc_blocks = (unsigned)((n_tiles + 255) / 256);
        kidz_crc_kernel<<<crc_blocks, 256, 0, st>>>(f->text.as<uint8_t>(), T, f->members.as<Member>(), (uint32_t)chain.members.size(),
                                                     f->consts.as<uint32_t>(), f->acc.as<uint32_t>());
        KID_COUNT_LAUNCH();
    }
    // newline counts per 4 KiB tile, their prefix sums in place, the total in misc[2]
    // misc (64-bit words): 0 non-empty lines, 1 records, 2 newlines, 3 error flags, 4 bases
    uint64_t *misc = f->misc.as<uint64_t>();
    KID_CUDA(f->tilecnt.reserve(sizeof(uint32_t) * ((size_t)n_tiles + 8)));
    uint32_t *tile_cnt = f->tilecnt.as<uint32_t>();
    if (n_tiles) {
        kidz_nl_count_kernel<<<n_tiles, 256, 0, st>>>(f->text.as<uint8_t>(), T, tile_cnt);
        KID_COUNT_LAUNCH();
    }
    KID_TRY(scan_u32<uint32_t>(f, tile_cnt, nullptr, n_tiles, tile_cnt, misc + 2));
    uint64_t *hs = f->h_small.as<uint64_t>();
    uint32_t *hcrc = reinterpret_cast<uint32_t *>(hs + 16);
    KID_CUDA(cudaMemcpyAsync(hs, misc, sizeof(uint64_t) * 16, cudaMemcpyDeviceToHost, st));
    if (!chain.members.empty())
        KID_CUDA(cudaMemcpyAsync(hcrc, f->acc.p, sizeof(uint32_t) * chain.members.size(), cudaMemcpyDeviceToHost, st));
    KID_CUDA(cudaStreamSynchronize(st));
    for (size_t i = 0; i < chain.members.size(); i++)
        if (hcrc[i] != chain.members[i].crc) return unsupported(path, "a member's CRC-32 does not match its trailer");
    if (hs[2] >= 0xfffffff0ull) return unsupported(path, "more than 2^32 lines");
    const uint32_t NL = (uint32_t)hs[2];
    double t5 = now_s();
    f->phase_s[kPhResolve] = t5 - t4;

    // ---- framing
    f->n_text = T;
    f->n_lines = NL;
    if (NL == 0 && T >= kRefLineLimit) return unsupported(path, "a line of 16 KiB or more");
    const size_t max_rec = (size_t)NL / 4 + 1;
    KID_CUDA(f->nlpos.reserve(sizeof(uint64_t) * ((size_t)NL + 8)));
    KID_CUDA(f->recs.reserve((3 * sizeof(uint64_t) + 3 * sizeof(uint32_t)) * max_rec));
    KID_CUDA(f->off.reserve(sizeof(uint64_t) * (max_rec + 1)));
    KID_CUDA(f->seq.reserve((size_t)(T / 2) + 256));
    KID_CUDA(f->qual.reserve((size_t)(T / 2) + 256));
    uint64_t *rec64 = f->recs.as<uint64_t>();
    uint32_t *rec32 = reinterpret_cast<uint32_t *>(rec64 + 3 * max_rec);
    f->rec = RecordArrays{ rec64, rec64 + max_rec, rec64 + 2 * max_rec, rec32, rec32 + max_rec, rec32 + 2 * max_rec };
    unsigned long long *err = reinterpret_cast<unsigned long long *>(misc + 3);
    const uint32_t line_blocks = (NL + 255) / 256;
    if (NL) {
        kidz_nl_write_kernel<<<n_tiles, 256, 0, st>>>(f->text.as<uint8_t>(), T, tile_cnt, f->nlpos.as<uint64_t>());
        KID_COUNT_LAUNCH();
        KID_CUDA(f->linecnt.reserve(sizeof(uint32_t) * ((size_t)line_blocks + 8)));
        uint32_t *line_cnt = f->linecnt.as<uint32_t>();
        kidz_line_count_kernel<<<line_blocks, 256, 0, st>>>(f->text.as<uint8_t>(), T, f->nlpos.as<uint64_t>(), NL, line_cnt, err);
        KID_COUNT_LAUNCH();
        KID_TRY(scan_u32<uint32_t>(f, line_cnt, nullptr, line_blocks, line_cnt, misc + 0));
        kidz_line_write_kernel<<<line_blocks, 256, 0, st>>>(f->text.as<uint8_t>(), f->nlpos.as<uint64_t>(), NL, line_cnt, misc, f->rec, err);
        KID_COUNT_LAUNCH();
    }
    kidz_set_records_kernel<<<1, 1, 0, st>>>(misc);
    KID_COUNT_LAUNCH();
    KID_TRY(scan_u32<uint64_t>(f, f->rec.seq_len, misc + 1, (uint32_t)max_rec, f->off.as<uint64_t>(), misc + 4));
    kidz_gather_kernel<<<(unsigned)f->db->sm_count * 8, 256, 0, st>>>(f->text.as<uint8_t>(), f->rec, misc, f->off.as<uint64_t>(),
                                                                       f->seq.as<uint8_t>(), f->qual.as<uint8_t>(), err);
    KID_COUNT_LAUNCH();
    KID_CUDA(cudaGetLastError());
    KID_CUDA(cudaMemcpyAsync(hs, misc, sizeof(uint64_t) * 16, cudaMemcpyDeviceToHost, st));
    KID_CUDA(cudaStreamSynchronize(st));
    if (hs[3] & kErrLongLine) return unsupported(path, "a line of 16 KiB or more");
    if (hs[3] & kErrShortQual) return unsupported(path, "a quality line shorter than its read");
    f->n_reads = (size_t)hs[1];
    f->n_bases = hs[4];
    if (kid_pack_word_index(f->n_bases, f->n_reads) + 2 >= 0x80000000ull) return unsupported(path, "more than 2^31 packed words");
    f->loaded = true;
    if (n_reads) *n_reads = f->n_reads;
    f->phase_s[kPhFrame] = now_s() - t5;
    return KID_OK;
}

int kid_fastq_classify(kid_fastq *f, kid_sample *s, int32_t *out_taxon)
{
    if (!f || !s) return kid_fail(KID_EINVAL, "kid_fastq_classify: NULL argument");
    if (!f->loaded) return kid_fail(KID_EINVAL, "kid_fastq_classify: no file is loaded");
    if (s->db != f->db) return kid_fail(KID_EINVAL, "kid_fastq_classify: the sample belongs to another database");
    if (s->db->layout != KID_LAYOUT_MINIMIZER) return kid_fail(KID_EINVAL, "kid_fastq_classify needs the default table layout");
    const size_t n = f->n_reads;
    if (n == 0) return KID_OK;
    DeviceGuard guard(f->db->device);
    const double t0 = now_s();
    cudaStream_t st = f->stream;
    const uint64_t bound = kid_pack_word_index(f->n_bases, n) + 2;
    if (bound >= 0x80000000ull) return kid_fail(KID_ERANGE, "kid_fastq_classify: %zu reads need more than 2^31 packed words", n);
    KID_CUDA(f->words.reserve(sizeof(uint32_t) * ((size_t)bound + 16)));
    KID_CUDA(f->meta.reserve(sizeof(uint2) * (n + 1)));
    KID_CUDA(f->taxon.reserve(sizeof(int32_t) * n));
    KID_CUDA(f->span.reserve(sizeof(uint32_t) * 2 * n));
    KID_CUDA(cudaStreamWaitEvent(st, s->begin_ev, 0)); // not before kid_sample_begin's memsets
    KidPackParams pp;
    pp.seq = f->seq.as<uint8_t>();
    pp.qual = f->qual.as<uint8_t>();
    pp.off = f->off.as<uint64_t>();
    pp.off_bias = 0;
    pp.n_reads = n;
    pp.words = f->words.as<uint32_t>();
    pp.meta = f->meta.as<uint2>();
    pp.out_span = f->span.as<uint32_t>();
    pp.accept_u = (s->db->flags & KID_DB_ACCEPT_U) != 0;
    KID_CUDA(kid_launch_pack(pp, s->db->sm_count, st));
    KidPackedParams p = kid_make_packed_params(s, f->words.as<uint32_t>(), f->meta.as<uint2>(), 0, n, f->taxon.as<int32_t>());
    KID_CUDA(kid_launch_classify3(p, s->db->sm_count, st));
    if (out_taxon) {
        KID_CUDA(cudaMemcpyAsync(out_taxon, f->taxon.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
        __atomic_fetch_add(&s->d2h, sizeof(int32_t) * n, __ATOMIC_RELAXED);
    }
    __atomic_fetch_add(&s->h2d, f->gz_bytes, __ATOMIC_RELAXED);
    KID_CUDA(cudaStreamSynchronize(st));
    f->phase_s[kPhClassify] = now_s() - t0;
    return KID_OK;
}

int kid_fastq_fetch(kid_fastq *f, const uint32_t *reads, size_t n, const char **data, const uint32_t **lens)
{
    if (!f || !data || !lens || (n && !reads)) return kid_fail(KID_EINVAL, "kid_fastq_fetch: NULL argument");
    *data = nullptr;
    *lens = nullptr;
    if (!f->loaded || !f->span.p) return kid_fail(KID_EINVAL, "kid_fastq_fetch: classify the file first");
    if (n == 0) return KID_OK;
    if (n >= 0x7fffffffull) return kid_fail(KID_ERANGE, "kid_fastq_fetch: too many reads");
    for (size_t i = 0; i < n; i++)
        if (reads[i] >= f->n_reads) return kid_fail(KID_ERANGE, "kid_fastq_fetch: read %u of %zu", reads[i], f->n_reads);
    DeviceGuard guard(f->db->device);
    const double t0 = now_s();
    cudaStream_t st = f->stream;
    KID_CUDA(f->fidx.reserve(sizeof(uint32_t) * n));
    KID_CUDA(f->flens.reserve(sizeof(uint32_t) * 2 * n));
    KID_CUDA(f->fpos.reserve(sizeof(uint32_t) * (n + 1)));
    KID_CUDA(f->h_fetch_lens.reserve(sizeof(uint32_t) * (2 * n + 1)));
    KID_CUDA(cudaMemcpyAsync(f->fidx.p, reads, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    kidz_fetch_plan_kernel<<<1, 1024, 0, st>>>(f->fidx.as<uint32_t>(), (uint32_t)n, f->rec, f->span.as<uint32_t>(), f->flens.as<uint32_t>(),
                                                f->fpos.as<uint32_t>());
    KID_COUNT_LAUNCH();
    KID_CUDA(cudaGetLastError());
    uint32_t *hl = f->h_fetch_lens.as<uint32_t>();
    KID_CUDA(cudaMemcpyAsync(hl, f->flens.p, sizeof(uint32_t) * 2 * n, cudaMemcpyDeviceToHost, st));
    KID_CUDA(cudaMemcpyAsync(hl + 2 * n, f->fpos.as<uint32_t>() + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    KID_CUDA(cudaStreamSynchronize(st));
    const size_t total = hl[2 * n];
    KID_CUDA(f->fdata.reserve(total + 16));
    KID_CUDA(f->h_fetch_data.reserve(total + 16));
    kidz_fetch_copy_kernel<<<(unsigned)std::min<size_t>((n + 7) / 8, (size_t)f->db->sm_count * 8), 256, 0, st>>>(
        f->text.as<uint8_t>(), f->fidx.as<uint32_t>(), (uint32_t)n, f->rec, f->span.as<uint32_t>(), f->flens.as<uint32_t>(),
        f->fpos.as<uint32_t>(), f->fdata.as<uint8_t>());
    KID_COUNT_LAUNCH();
    KID_CUDA(cudaGetLastError());
    if (total) KID_CUDA(cudaMemcpyAsync(f->h_fetch_data.p, f->fdata.p, total, cudaMemcpyDeviceToHost, st));
    KID_CUDA(cudaStreamSynchronize(st));
    *data = f->h_fetch_data.as<char>();
    *lens = hl;
    f->phase_s[kPhFetch] += now_s() - t0;
    return KID_OK;
}

int kid_fastq_stats(const kid_fastq *f, uint64_t *n_text, uint64_t *n_pieces, uint64_t *n_again, uint64_t *n_members, double *phase_seconds,
                    int n_phases)
{
    if (!f) return kid_fail(KID_EINVAL, "kid_fastq_stats: f is NULL");
    if (n_text) *n_text = f->n_text;
    if (n_pieces) *n_pieces = f->n_pieces;
    if (n_again) *n_again = f->n_redo;
    if (n_members) *n_members = f->n_members;
    for (int i = 0; phase_seconds && i < n_phases; i++) phase_seconds[i] = i < kPhCount ? f->phase_s[i] : 0.0;
    return KID_OK;
}

} // extern "C"
