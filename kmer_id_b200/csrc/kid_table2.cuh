// kid_table2.cuh - the minimizer-addressed probe table ("layout M", the default).
//
// Why: ncu on the key-hashed table (kid_common.cuh, "layout K") shows every lookup pulling its own
// 128-byte line from DRAM (profiles/r1_v1_keyhash_classify_ncu.txt: 4 DRAM sectors per lookup), and
// the chip serves only ~37.8 G such independent requests per second whatever their size
// (profiles/r1_gather_microbench_ncu.txt).  One request per lookup therefore caps the path at
// ~38 G lookups/s.  The only way up is to make consecutive k-mers of a read share requests.
//
// How: the home of a key is chosen by its MINIMIZER, not by the key itself.
//   c(i)     = kid_mm_hash(canonical 16-mer starting at base i)        i = 0..14 within the 30-mer
//   M(key)   = min_i c(i)                                              strand independent
//   line     = (M * 0x9E3779B1) >> (32 - L)                            2^L lines of 128 bytes
//   sector   = 4 * line + top 2 bits of kid_key_hash32(key)            4 sectors per line
//   sector   = 3 entries, split so that the first 16 bytes decide almost every lookup:
//                words 0..2  high words of the three keys: (key >> 32) | 1<<31, 0 = empty slot
//                word  3     low 32 bits of the three 21-bit taxa
//                words 4..6  low words of the three keys
//                word  7     high 31 bits of the taxa
//              12 entries per 128-byte line.  A lookup reads the first half (4 registers) and compares
//              the 29-bit high words; only a lane whose high word matches (a hit, or one miss in 2^28)
//              reads the second half - from L1, where its sector just arrived - to compare the low
//              word and pick up the taxon.  Four chunks of lookups fit the registers that two took
//              when whole sectors were loaded.
// Adjacent k-mers of a read share their minimizer in runs of ~7.5, so the 32 lanes of a warp (32
// consecutive k-mers) touch ~5 distinct lines instead of 32: ~6.5x fewer DRAM line fetches and
// L2 requests per lookup.  Within the line the key picks the sector, so one lane still reads just
// 32 bytes, and a genome region whose k-mers share a minimizer spreads over the line's 12 slots.
// A key that finds its sector full lives in a following sector (linear probing in sector units,
// built in sorted order by kid_build_sorted.cu, slack sectors instead of wrap-around); every slot
// between a key's home and its slot is occupied, so a lookup stops at the first sector that has
// an empty slot.
// Full 60-bit keys are stored, so a match is exact by construction (no fingerprints).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define KID_MM 16                       /* minimizer length in bases */
#define KID_MM_WINDOWS (30 - KID_MM + 1) /* 15 candidate 16-mers per 30-mer */
#define KID2_OCC (1ULL << 63)
#define KID2_MIN_LOG2_LINES 10
#define KID2_MAX_LOG2_LINES 30 /* sector indices stay below 2^32 */

#define KID2_SLOTS_PER_SECTOR 3
#define KID2_TAXON_BITS 21
#define KID2_MAX_TAXA ((1 << KID2_TAXON_BITS) - 1)

struct Kid2TableView {
    const uint4 *sectors; // 2 uint4 (= 32 bytes, 3 entries) per sector
    uint64_t sector_mask; // n_sectors - 1 (home sectors; the table has slack sectors after them)
    int line_shift;       // 32 - log2(groups): the minimizer picks a group of 2^sub_bits sectors
    int max_probe;        // longest displacement (in sectors) any key needed at build time
    int sub_bits;         // 2 = one 128-byte line per minimizer (default); 3, 4 for very large DBs
    int mm;               // minimizer length in bases: 16 (default) or 20 (databases beyond ~4e8 keys)
};

// reverse complement of a 16-mer held in 32 bits (first base in the top pair)
__host__ __device__ __forceinline__ uint32_t kid_rc16(uint32_t x)
{
#ifdef __CUDA_ARCH__
    uint32_t y = __brev(x);
#else
    uint32_t y = x;
    y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
    y = ((y >> 2) & 0x33333333u) | ((y & 0x33333333u) << 2);
    y = ((y >> 4) & 0x0F0F0F0Fu) | ((y & 0x0F0F0F0Fu) << 4);
    y = ((y >> 8) & 0x00FF00FFu) | ((y & 0x00FF00FFu) << 8);
    y = (y >> 16) | (y << 16);
#endif
#ifdef __CUDA_ARCH__
    // ~(((y >> 1) & 0x55555555) | ((y << 1) & 0xAAAAAAAA)) as ONE lop3 (LUT 0x1B = ~((a&c)|(b&~c)))
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, 0x55555555, 0x1B;" : "=r"(r) : "r"(y >> 1), "r"(y << 1));
    return r;
#else
    y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
    return ~y;
#endif
}

// ordering hash of a canonical 16-mer (only its order matters; it need not be a bijection)
__host__ __device__ __forceinline__ uint32_t kid_mm_hash_canon(uint32_t x)
{
    // One multiply: the order is decided by the high bits of the product, which depend on every bit of
    // x.  The xor keeps x = 0 (poly-A / poly-T) from being the smallest value of all, i.e. from winning
    // every window it occurs in.  Measured against multiply - xorshift - multiply (tools/gpu_runs/gpu_m.sh):
    // fewer displaced keys (83 k vs 98 k at bact10 scale, 19.5 M vs 24.9 M at 10x) and 1-4 % more lookups/s.
    return (x ^ 0x5BD1E995u) * 0x9E3779B1u;
}
__host__ __device__ __forceinline__ uint32_t kid_mm_hash(uint32_t fwd16)
{
    const uint32_t rc = kid_rc16(fwd16);
    return kid_mm_hash_canon(fwd16 < rc ? fwd16 : rc);
}

// minimizer hash of a 60-bit key (build side; the classify kernel slides it along the read)
__host__ __device__ __forceinline__ uint32_t kid_minimizer(uint64_t key)
{
    uint32_t m = 0xFFFFFFFFu;
    for (int i = 0; i < KID_MM_WINDOWS; i++) {
        const uint32_t c = kid_mm_hash((uint32_t)(key >> (2 * (KID_MM_WINDOWS - 1 - i))));
        m = c < m ? c : m;
    }
    return m;
}

// ---- m = 20 for very large databases (KID_DB_MM=20) ------------------------------------------------
// Only ~10 % of the 2^31 canonical 16-mers ever win a 15-way minimum, so beyond a few 1e8 keys several
// keys share every minimizer and pile into its 12-entry line.  A 20-mer minimizer (11 windows, 2^39
// canonical values) gives practically every key of a 1e9-key database its own - but a 32-bit word cannot
// say which: the minimum of 11 hashes lies in the lowest twelfth of the range, so 1e9 winners share
// ~4e8 values whatever the hash.  Each candidate therefore carries 64 bits, (order hash, identity hash),
// the sliding minimum runs over the pair, and the winner's IDENTITY word addresses the line.
// The canonical 20-mer is handled as (top 32 bits, low 8 bits).
#define KID_MM20 20
#define KID_MM20_WINDOWS (30 - KID_MM20 + 1) /* 11 */
__host__ __device__ __forceinline__ uint64_t kid_mm20_pair(uint32_t top, uint32_t low8)
{
    const uint32_t order = (top ^ 0x5BD1E995u) * 0x9E3779B1u ^ low8 * 0xC2B2AE3Du;
    const uint32_t ident = (top ^ 0x7F4A7C15u) * 0x85EBCA77u + low8 * 0x27D4EB2Fu;
    return ((uint64_t)order << 32) | ident;
}
// what addresses the line of a 60-bit key with 20-mer minimizers (build side): the identity word of the
// smallest (order, identity) pair among its 11 candidates
__host__ __device__ __forceinline__ uint32_t kid_minimizer20(uint64_t key)
{
    uint64_t m = ~0ull;
    for (int i = 0; i < KID_MM20_WINDOWS; i++) {
        const uint64_t f = (key >> (2 * (KID_MM20_WINDOWS - 1 - i))) & ((1ULL << 40) - 1ULL); // bases i..i+19
        // reverse complement of 20 bases: that of the 16 first bases below that of the last 4
        const uint64_t r = ((uint64_t)(kid_rc16((uint32_t)(f << 24)) & 0xFFu) << 32) | kid_rc16((uint32_t)(f >> 8));
        const uint64_t c = f < r ? f : r;
        const uint64_t h = kid_mm20_pair((uint32_t)(c >> 8), (uint32_t)(c & 0xFFu));
        m = h < m ? h : m;
    }
    return (uint32_t)m;
}
__host__ __device__ __forceinline__ uint32_t kid_minimizer_mm(uint64_t key, int mm)
{
    return mm == KID_MM20 ? kid_minimizer20(key) : kid_minimizer(key);
}

// the key's own hash picks the sector inside the minimizer's group: its top sub_bits bits
__host__ __device__ __forceinline__ uint32_t kid_key_hash32(uint64_t key)
{
    return ((uint32_t)key ^ (uint32_t)(key >> 32)) * 0xC2B2AE35u;
}

// sub_bits = log2(sectors per minimizer-addressed group).  2 = one 128-byte line, which is what keeps
// neighbouring k-mers of a read on one DRAM line.  With m = 16 only ~10 % of the 2^31 canonical
// 16-mers ever win a 15-way minimum, so beyond ~3e8 keys several keys per minimizer pile into one
// 12-entry line; 3 or 4 spreads a minimizer over 2 or 4 lines instead.
__host__ __device__ __forceinline__ uint64_t kid2_home_sector(uint32_t minimizer, uint64_t key,
                                                              int line_shift, int sub_bits)
{
    const uint32_t grp = line_shift >= 32 ? 0u : (minimizer * 0x9E3779B1u) >> line_shift;
    return ((uint64_t)grp << sub_bits) | (kid_key_hash32(key) >> (32 - sub_bits));
}

#ifdef __CUDACC__

// first half of a sector (high words + taxa low word); the sector stays in L1 for the second half
__device__ __forceinline__ uint4 kid2_load_half(const uint4 *p)
{
    uint4 a;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(p));
    return a;
}
// same, but only lanes with pred != 0 issue the load (a keeps its value otherwise): no branch
__device__ __forceinline__ void kid2_load_half_if(const uint4 *p, uint4 &a, uint32_t pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t"
                 "@q ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
                 : "+r"(a.x), "+r"(a.y), "+r"(a.z), "+r"(a.w)
                 : "l"(p), "r"(pred));
}

// taxon of entry j (0..2): three 21-bit fields in words 3 (low) and 7 (high)
__host__ __device__ __forceinline__ uint32_t kid2_taxon_of(uint32_t w3, uint32_t w7, int j)
{
    const uint64_t t = ((uint64_t)w7 << 32) | w3;
    return (uint32_t)(t >> (KID2_TAXON_BITS * j)) & (uint32_t)KID2_MAX_TAXA;
}

// which entries of a first half carry this high word (bit j per entry)
__device__ __forceinline__ uint32_t kid2_candidates(const uint4 &h, uint32_t want_hi)
{
    return (h.x == want_hi ? 1u : 0u) | (h.y == want_hi ? 2u : 0u) | (h.z == want_hi ? 4u : 0u);
}
// all three slots occupied (occupied high words carry bit 31)
__device__ __forceinline__ bool kid2_full(const uint4 &h) { return (int32_t)(h.x & h.y & h.z) < 0; }

// Exact match of one sector given its first half and the candidate mask: reads the second half.
// 1 = hit (taxon, slot j), 0 = no entry holds the key.
__device__ __forceinline__ int kid2_verify(const uint4 *sector, const uint4 &h, uint32_t cand, uint32_t want_lo,
                                           uint32_t &taxon, int &j)
{
    const uint4 l = kid2_load_half(sector + 1);
    const uint32_t ok = cand & ((l.x == want_lo ? 1u : 0u) | (l.y == want_lo ? 2u : 0u) | (l.z == want_lo ? 4u : 0u));
    if (!ok) return 0;
    j = __ffs(ok) - 1;
    taxon = kid2_taxon_of(h.w, l.w, j);
    return 1;
}

// continue a lookup from sector `s` (used for the rare full sectors and by the diagnostic kernel)
__device__ __forceinline__ uint32_t kid2_lookup_from(const Kid2TableView &t, uint64_t s, uint64_t key,
                                                     int probes_done, uint64_t &slot)
{
    const uint32_t lo = (uint32_t)key, hi = (uint32_t)(key >> 32) | 0x80000000u;
    for (int d = probes_done; d <= t.max_probe; d++) {
        const uint64_t sec = s + (uint64_t)d; // clusters run into the slack sectors, never wrap
        const uint4 h = kid2_load_half(t.sectors + 2 * sec);
        const uint32_t cand = kid2_candidates(h, hi);
        if (cand) {
            uint32_t taxon;
            int j;
            if (kid2_verify(t.sectors + 2 * sec, h, cand, lo, taxon, j)) {
                slot = KID2_SLOTS_PER_SECTOR * sec + (uint64_t)j;
                return taxon;
            }
        }
        if (!kid2_full(h)) return 0; // an empty slot ends the cluster
    }
    return 0;
}

#endif // __CUDACC__
