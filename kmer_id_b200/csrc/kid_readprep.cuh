// kid_readprep.cuh - device routines that turn raw read bytes into what the k-mer scan consumes:
// the exact ACGTacgt(+Uu) test with 2-bit packing, and process_qual's trim (newkmer_10nx.cpp:714-760).
// Used by the batch packer (kid_pack.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// 4 ASCII bases in one 32-bit word (first base in the low byte) -> 8 bits of 2-bit codes with the
// first base in the top pair, and a 4-bit validity mask with the first base in the top bit.
// Bits 2..1 of the byte are the raw code r (A 0, C 1, T/U 2, G 3); ignoring those and the case bit,
// an A/C/G byte equals 0x41 and a T byte equals 0x41 ^ 0x11, so one xor/and and an exact
// zero-byte test decide ACGTacgt (+Uu) for four bases at once.
__device__ __forceinline__ void pack4(uint32_t x, bool accept_u, uint32_t &code8, uint32_t &valid4)
{
    const uint32_t s1 = x >> 1, s2 = x >> 2;
    const uint32_t tflag = s2 & ~s1 & 0x01010101u; // r == 2
    uint32_t z = ((x ^ 0x41414141u) & 0xD9D9D9D9u) ^ (tflag * 0x11u);
    if (accept_u) z &= ~tflag; // 'U' differs from 'T' in bit 0 only (kmer_read_vf6.cpp:496-500)
    const uint32_t nz = (((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z) & 0x80808080u; // 0x80 per non-zero byte
    valid4 = ((((nz ^ 0x80808080u) >> 7) * 0x08040201u) >> 24) & 0xFu;
    uint32_t c = s1 & 0x03030303u; // swap 2<->3 -> A0 C1 G2 T3 (:480-519)
    c ^= (c >> 1) & 0x01010101u;
    code8 = (c * 0x40100401u) >> 24;
}

// generic 32-positions-per-step scans (fallbacks of the trim fast path) -----------------------
__device__ __forceinline__ int scan_first_good(const signed char *q, int from, int stop, int lane)
{ // first i in [from, stop) with q[i] >= '1', else stop
    for (int base = from; base < stop; base += 32) {
        const int i = base + lane;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, i < stop && q[i] >= 49);
        if (m) return base + __ffs(m) - 1;
    }
    return stop;
}
__device__ __forceinline__ int scan_last_good(const signed char *q, int from, int start, int lane)
{ // last i in (start, from] with q[i] >= '1', else start
    for (int hi = from; hi > start; hi -= 32) {
        const int i = hi - lane;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, i > start && q[i] >= 49);
        if (m) return hi - (__ffs(m) - 1);
    }
    return start;
}
__device__ __forceinline__ int scan_window_fwd(const signed char *q, int start, int lim, int lane)
{ // first s in [start, lim) whose 4-window sum(q-32) >= 68, else lim
    for (int base = start; base < lim; base += 32) {
        const int s = base + lane;
        bool ok = false;
        if (s < lim) ok = (int)q[s] + q[s + 1] + q[s + 2] + q[s + 3] - 128 >= 68;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
        if (m) return base + __ffs(m) - 1;
    }
    return lim;
}
__device__ __forceinline__ int scan_window_bwd(const signed char *q, int stop, int lo, int lane)
{ // last t in (lo, stop] whose trailing 4-window sum >= 68, else lo
    for (int hi = stop; hi > lo; hi -= 32) {
        const int t = hi - lane;
        bool ok = false;
        if (t > lo) ok = (int)q[t] + q[t - 1] + q[t - 2] + q[t - 3] - 128 >= 68;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
        if (m) return hi - (__ffs(m) - 1);
    }
    return lo;
}

// The read stream is used once: keep it out of L1, which then holds the taxonomy rows and the spill
// slots (+0.9 % lookups/s measured).
__device__ __forceinline__ uint4 load_stream16(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// ---- TRIM (:724-753) from the two preloaded quality registers qa = q[lane], qb = q[len-1-lane]
__device__ __forceinline__ void trim_read(const signed char *q, int len, int qa, int qb, int lane, int &start,
                                          int &stop)
{
    const unsigned full = 0xFFFFFFFFu;
    start = 0;
    stop = len - 1;
    if (len <= 0) return;
    // fast path (most reads): both end bases and both end windows pass -> nothing to trim
    if (len >= 6) {
        int sa = qa + __shfl_down_sync(full, qa, 1), sb = qb + __shfl_down_sync(full, qb, 1);
        sa += __shfl_down_sync(full, sa, 2); // lane 0: q[0]+q[1]+q[2]+q[3]
        sb += __shfl_down_sync(full, sb, 2); // lane 0: q[len-1]+...+q[len-4]
        if (__shfl_sync(full, (int)(qa >= 49 && qb >= 49 && sa - 128 >= 68 && sb - 128 >= 68), 0)) return;
    }
    { // while (qual[start] < '1' && start < stop) start++;
        const unsigned m = __ballot_sync(full, lane < stop && qa >= 49);
        start = m ? __ffs(m) - 1 : scan_first_good(q, 32, stop, lane);
    }
    { // while (qual[stop] < '1' && stop > start) stop--;
        const unsigned m = __ballot_sync(full, len - 1 - lane > start && qb >= 49);
        stop = m ? len - 1 - (__ffs(m) - 1) : scan_last_good(q, len - 33, start, lane);
    }
    if (start < stop - 4) { // leading 4-base window slides right while sum(q-32) < 68
        const int lim = stop - 4, s = start + lane;
        const int w = __shfl_sync(full, qa, s & 31) + __shfl_sync(full, qa, (s + 1) & 31) +
                      __shfl_sync(full, qa, (s + 2) & 31) + __shfl_sync(full, qa, (s + 3) & 31);
        const bool known = s + 3 < 32;
        const unsigned mk = __ballot_sync(full, known && s < lim && w - 128 >= 68);
        const unsigned unk = __ballot_sync(full, !known && s < lim);
        if (mk && (!unk || __ffs(mk) < __ffs(unk))) start += __ffs(mk) - 1;
        else if (!mk && !unk) start = lim;
        else start = scan_window_fwd(q, start, lim, lane);
    }
    if (start < stop - 4) { // trailing window slides left
        const int lo = start + 4, t = stop - lane;
        const int idx = len - 1 - t; // lane of qb that holds q[t]
        const int w = __shfl_sync(full, qb, idx & 31) + __shfl_sync(full, qb, (idx + 1) & 31) +
                      __shfl_sync(full, qb, (idx + 2) & 31) + __shfl_sync(full, qb, (idx + 3) & 31);
        const bool known = idx + 3 < 32;
        const unsigned mk = __ballot_sync(full, known && t > lo && w - 128 >= 68);
        const unsigned unk = __ballot_sync(full, !known && t > lo);
        if (mk && (!unk || __ffs(mk) < __ffs(unk))) stop -= __ffs(mk) - 1;
        else if (!mk && !unk) stop = lo;
        else stop = scan_window_bwd(q, stop, lo, lane);
    }
}

// ---- TRIM, one lane per read: process_qual's four loops (:727-753) as they stand, for the ordinary case
// of a short low-quality head or tail.  Returns false when `budget` steps were not enough (a long run of
// bad qualities): the caller then hands the read to the warp-cooperative searches above.
__device__ __forceinline__ bool trim_read_lane(const signed char *q, int len, int budget, int &start, int &stop)
{
    start = 0;
    stop = len - 1;
    if (len <= 0) return true;
    while (start < stop && q[start] < 49) { // :727-728
        start++;
        if (--budget < 0) return false;
    }
    while (stop > start && q[stop] < 49) { // :729-730
        stop--;
        if (--budget < 0) return false;
    }
    if (start < stop - 4) { // :732-742 leading 4-base window
        int w = q[start] + q[start + 1] + q[start + 2] + q[start + 3] - 128;
        while (w < 68 && start < stop - 4) {
            w += q[start + 4] - q[start];
            start++;
            if (--budget < 0) return false;
        }
    }
    if (start < stop - 4) { // :743-753 trailing window
        int w = q[stop] + q[stop - 1] + q[stop - 2] + q[stop - 3] - 128;
        while (w < 68 && start < stop - 4) {
            w += q[stop - 4] - q[stop];
            stop--;
            if (--budget < 0) return false;
        }
    }
    return true;
}

