// kid_common.cuh - table layout, hash, device lookup and taxonomy routines shared by the kernels.
//
// GPU-resident probe table (replaces Hashtable, newkmer_10nx.cpp:158-265)
// ------------------------------------------------------------------------
// The reference keeps 2^30 24-byte cells and walks a triangular probe sequence, one dependent
// DRAM access per step.  Only key -> taxon survives on this path (org/position/fstrand feed the
// dead Smith-Waterman branch), so the device table is built for ONE 32-byte DRAM sector per
// lookup:
//
//   h       = kid_hash60(key)            a bijection on 60-bit keys
//   bucket  = h >> (60 - B)              B = log2(#buckets), 22 <= B <= 32
//   rem     = h & (2^(60-B) - 1)         <= 38 bits: with the bucket index it identifies the key
//   entry   = rem << 26 | disp << 22 | taxon      one uint64; 0 = empty (taxon 0 is never stored)
//   bucket  = 4 entries = 32 bytes, 32-byte aligned  -> one sector, one LDG.256
//
// A key that does not fit its home bucket lives in bucket home+disp (disp <= 15, stored so that the
// remainder still identifies the key; placement by kid_build_sorted.cu, slack buckets instead of
// wrap-around).  Every slot between a key's home and its slot is occupied, so a lookup may stop at
// the first bucket that has an empty slot.  At the default load (<= 1 key per bucket on average)
// a second sector is needed by ~1 % of lookups.  This is "layout K", kept for the bake-off against
// the minimizer-addressed layout M (kid_table2.cuh), which is the default.
//
// "seen" flags (the reference's std::set kmer_seen, :64,:596-603) are one bit per slot in a
// separate per-sample bitmap, so the table itself is read-only on the hot path.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define KID_TAXON_BITS 22
#define KID_DISP_BITS 4
#define KID_TAXON_MASK ((1u << KID_TAXON_BITS) - 1u)
#define KID_MAX_DISP ((1 << KID_DISP_BITS) - 1)
#define KID_MAX_TAXA ((int)KID_TAXON_MASK)
#define KID_MIN_LOG2_BUCKETS 22
#define KID_MAX_LOG2_BUCKETS 32
#define KID_MASK60 ((1ULL << 60) - 1ULL)

struct KidTableView {
    const uint64_t *slots; // 4 * n_buckets entries
    uint64_t bucket_mask;  // n_buckets - 1
    int rem_bits;          // 60 - log2_buckets
};

// {parent, depth}; parent[0] = parent[1] = 1 (Tree1::get_parent :146-152), depth[1] = 0
struct KidTreeView {
    const uint2 *node;
    int n_taxa;
};

// Bijective mixer on [0, 2^60): xor-shifts and odd multipliers mod 2^60 are each invertible.
__host__ __device__ __forceinline__ uint64_t kid_hash60(uint64_t k)
{
    k ^= k >> 31;
    k = (k * 0x9E3779B97F4A7C15ULL) & KID_MASK60;
    k ^= k >> 29;
    k = (k * 0xBF58476D1CE4E5B9ULL) & KID_MASK60;
    k ^= k >> 32;
    return k;
}

// Reverse complement of a 60-bit forward key: identical to the rolling keyR of
// newkmer_10nx.cpp:482-517 once 30 bases are in (base i of the window sits at bits 2i).
__host__ __device__ __forceinline__ uint64_t kid_revcomp60(uint64_t kf)
{
#ifdef __CUDA_ARCH__
    uint64_t x = __brevll(kf);
#else
    uint64_t x = kf;
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
    x = (x >> 32) | (x << 32);
#endif
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    return (~(x >> 4)) & KID_MASK60;
}

#ifdef __CUDACC__

__device__ __forceinline__ void kid_load_bucket(const uint64_t *p, uint64_t e[4])
{
    // one 32-byte sector; read-only for the whole kernel, no reuse -> keep it out of L1
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(e[0]), "=l"(e[1]), "=l"(e[2]), "=l"(e[3])
                 : "l"(p));
}

// Compare one loaded bucket against (rem, disp).  Returns 1 = hit (taxon/slot set), 0 = miss is
// final (bucket has an empty slot), -1 = bucket full without a match: continue with disp + 1.
__device__ __forceinline__ int kid_match_bucket(const uint64_t e[4], uint64_t tag, uint32_t &taxon,
                                                int &slot_in_bucket)
{
    bool has_empty = false;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if ((e[j] >> KID_TAXON_BITS) == tag) {
            taxon = (uint32_t)e[j] & KID_TAXON_MASK;
            slot_in_bucket = j;
            return 1;
        }
        has_empty |= (e[j] == 0);
    }
    return has_empty ? 0 : -1;
}

// Hashtable::getHash (:204-233) on the device table.  slot = global slot index of the hit.
__device__ __forceinline__ uint32_t kid_lookup_from(const KidTableView &t, uint64_t h, int disp0,
                                                    uint64_t &slot)
{
    const uint64_t home = h >> t.rem_bits;
    const uint64_t rem = h & ((1ULL << t.rem_bits) - 1ULL);
    for (int d = disp0; d <= KID_MAX_DISP; d++) {
        const uint64_t b = home + (uint64_t)d; // slack sectors after the last home bucket, no wrap
        uint64_t e[4];
        kid_load_bucket(t.slots + 4 * b, e);
        uint32_t taxon;
        int j;
        int r = kid_match_bucket(e, (rem << KID_DISP_BITS) | (uint64_t)d, taxon, j);
        if (r > 0) { slot = 4 * b + (uint64_t)j; return taxon; }
        if (r == 0) return 0;
    }
    return 0;
}

// Tree1::msca (:118-144): x if y is root or an ancestor-or-self of x; y if x is an ancestor of y;
// otherwise their lowest common ancestor.  Restated with depths instead of a std::set.
__device__ __forceinline__ uint32_t kid_msca(const KidTreeView &tr, uint32_t x, uint32_t y)
{
    if (x == y) return x;
    uint32_t a = x, b = y;
    uint2 na = __ldg(tr.node + a), nb = __ldg(tr.node + b);
    uint32_t da = na.y, db = nb.y;
    while (da > db) { a = na.x; na = __ldg(tr.node + a); da--; }
    while (db > da) { b = nb.x; nb = __ldg(tr.node + b); db--; }
    while (a != b) {
        a = na.x; b = nb.x;
        na = __ldg(tr.node + a);
        nb = __ldg(tr.node + b);
    }
    if (a == y) return x;
    if (a == x) return y;
    return a;
}

#endif // __CUDACC__
