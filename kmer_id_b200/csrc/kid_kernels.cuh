// kid_kernels.cuh - launch interfaces between kid_api.cu and the kernel files.
#pragma once
#include "kid_common.cuh"
#include "kid_table2.cuh"

#define KID_KSIZE 30
// Layout M scan kernel: one 1024-thread block per SM at 64 registers per thread.  Measured on B200
// (round 1, same 32 resident warps per SM): 4 x 256 threads 99.7 G lookups/s, 2 x 512 104.7 G,
// 1 x 1024 112.3 G (768 threads at 80 registers: 97.2 G) - one shared-memory gcount histogram per SM
// instead of four leaves more of the 256 KB array to L1.
#ifndef KID_CLASSIFY_THREADS
#define KID_CLASSIFY_THREADS 1024
#endif
#define KID_CLASSIFY1_THREADS 256 // layout K kernel (bake-off only)
#define KID_SMEM_HIST_MAX_BYTES (96u * 1024u) /* gcount histogram lives in smem up to 24576 taxa */

enum KidLayout { KID_LAYOUT_MINIMIZER = 0, KID_LAYOUT_KEYHASH = 1 };

struct KidClassifyParams {
    KidTableView table;   // layout K (kid_common.cuh)
    Kid2TableView table2; // layout M (kid_table2.cuh)
    KidTreeView tree;
    const uint8_t *seq;
    const uint8_t *qual; // NULL: no trimming
    const uint64_t *off; // n_reads + 1
    uint64_t off_bias;   // subtracted from every off[] value (chunked host batches)
    size_t n_reads;
    int32_t *out_taxon;  // may be NULL
    uint32_t *out_span;  // may be NULL
    int *gcount;
    uint32_t *seen;
    unsigned long long *counters; // [0] lookups, [1] hits
    bool accept_u;
};

// ---- packed read batches --------------------------------------------------------------------------
// What the k-mer scan consumes (include/kmer_id.h, "packed read batches"): per read the TRIMMED bases
// as 2-bit codes, 16 per 32-bit word with the first base in the top pair (A 0, C 1, G 2, T 3, anything
// else 0), zero padded to a whole word; reads that contain a non-ACGT base carry KID_PK_FLAG and are
// followed by ceil(tlen/32) validity words (first base in the top bit, 1 = ACGT).
//   meta[r] = { first word of read r | KID_PK_FLAG, tlen }     r = 0..n_reads (entry n_reads = end)
// tlen = stop - start + 1 of process_qual (newkmer_10nx.cpp:714-760); a read is classified iff
// tlen >= 31 (:755).  Words need not be dense: kid_pack.cu leaves gaps, a host packer does not.
#define KID_PK_FLAG 0x80000000u
struct KidPackedParams {
    Kid2TableView table2;
    KidTreeView tree;
    const uint32_t *words;
    const uint2 *meta;  // n_reads + 1
    uint32_t word_bias; // subtracted from every meta[].x word index (chunked host batches)
    size_t n_reads;
    int32_t *out_taxon; // may be NULL
    int *gcount;
    uint32_t *seen;
    unsigned long long *counters; // [0] lookups, [1] hits
};
cudaError_t kid_launch_classify3(const KidPackedParams &p, int sm_count, cudaStream_t stream);

// raw bytes -> packed batch on the device (kid_pack.cu): process_qual's trim + ACGT test + 2-bit pack.
// Read r gets the words starting at (off[r]-bias)/16 + (off[r]-bias)/32 + 2r (room for its validity
// words whether it needs them or not), so no prefix sum is needed.
struct KidPackParams {
    const uint8_t *seq;
    const uint8_t *qual; // NULL: no trimming (FASTA)
    const uint64_t *off; // n_reads + 1
    uint64_t off_bias;
    size_t n_reads;
    uint32_t *words;     // kid_pack_words_bound(bytes, n_reads) entries
    uint2 *meta;         // n_reads + 1
    uint32_t *out_span;  // may be NULL: {start, stop} per read as process_qual leaves them
    bool accept_u;
};
__host__ __device__ __forceinline__ uint64_t kid_pack_word_index(uint64_t rel_off, uint64_t r)
{
    return (rel_off >> 4) + (rel_off >> 5) + 2 * r;
}
cudaError_t kid_launch_pack(const KidPackParams &p, int sm_count, cudaStream_t stream);

// dense batch (include/kmer_id.h) -> packed batch, same word placement as kid_pack_kernel
struct KidExpandParams {
    const uint32_t *codes;    // codes[0] = the 16 bases from stream position `bias` on
    const uint32_t *boff;     // n_reads + 1 stream positions
    uint32_t bias;            // multiple of 16
    const uint32_t *flagbits; // bit r % 32 of word r / 32 (r = 0 is this batch's first read)
    const uint32_t *inv;      // stream positions of the non-ACGT bases of these reads, ascending
    uint32_t n_inv;
    size_t n_reads;
    uint32_t *words;
    uint2 *meta; // n_reads + 1
};
cudaError_t kid_launch_expand(const KidExpandParams &p, int sm_count, cudaStream_t stream);

// every kernel launch of this library bumps this (bench.py reports it as gpu_launches)
extern unsigned long long g_kid_kernel_launches;
#define KID_COUNT_LAUNCH() (__atomic_add_fetch(&g_kid_kernel_launches, 1ULL, __ATOMIC_RELAXED))

cudaError_t kid_launch_classify(const KidClassifyParams &p, int sm_count, cudaStream_t stream);  // layout K

// ---- table build (kid_build_sorted.cu) ----------------------------------------------------------
struct Kid2BuildStatus {
    unsigned long long n_distinct;
    unsigned long long n_displaced; // keys outside their home sector
    unsigned int max_probe;         // longest displacement in sectors
    unsigned int range_error;       // some taxa[i] >= n_taxa (or too wide for the layout)
    unsigned int overflow;          // a key fell off the slack or exceeded max_disp
    unsigned int pad;
};
struct KidSortedBuildParams {
    int layout;            // KidLayout
    int slots_per_sector;  // 4 (K) or 3 (M)
    uint64_t n_sectors;    // addressable home sectors (power of two)
    uint64_t slack_sectors; // extra sectors after the last home sector (no wrap-around)
    int line_shift;        // layout M: 32 - log2(groups)
    int sub_bits;          // layout M: log2(sectors per minimizer-addressed group), 2..4
    int mm;                // layout M: minimizer length, 16 or 20
    int rem_bits;          // layout K: 60 - log2_sectors
    uint32_t n_taxa, max_taxon, max_disp;
};
#define KID2_SLACK_SECTORS 4096
#define KID1_SLACK_SECTORS 16
// owner: uint32[(n_sectors + slack) * slots_per_sector]; synchronises `stream`
cudaError_t kid_build_owner_sorted(const uint64_t *keys, const uint32_t *taxa, size_t n, const KidSortedBuildParams &p,
                                   uint32_t *owner, Kid2BuildStatus *dstatus, Kid2BuildStatus *hstatus,
                                   cudaStream_t stream);
cudaError_t kid_launch_pack2(uint4 *sectors, size_t n_sectors_total, const uint32_t *owner, const uint64_t *keys,
                             const uint32_t *taxa, cudaStream_t stream);
cudaError_t kid_launch_pack1(uint64_t *slots, size_t n_slots_total, int rem_bits, const uint32_t *owner,
                             const uint64_t *keys, const uint32_t *taxa, cudaStream_t stream);
cudaError_t kid_launch_lookup2(const Kid2TableView &t, const uint64_t *keys, size_t n, uint32_t *out,
                               cudaStream_t stream);

// ---- sample-end kernels (kid_sample.cu) ------------------------------------------------------------
// slots: layout K = uint64 entries, layout M = packed 32-byte sectors of 3 entries
cudaError_t kid_launch_ucount(const void *slots, int layout, const uint32_t *seen, uint64_t word0,
                              uint64_t n_words, int *ucount, int n_taxa, cudaStream_t stream);
#define KID_MAX_OR_SOURCES 16
struct KidPtrList {
    const uint32_t *p[KID_MAX_OR_SOURCES];
};
// fused OR over n_src bitmaps + histogram (sources may be peer-mapped)
cudaError_t kid_launch_ucount_or(const void *slots, int layout, const KidPtrList &src, int n_src, uint64_t word0,
                                 uint64_t n_words, int *ucount, int n_taxa, cudaStream_t stream);
cudaError_t kid_launch_seen_or(uint32_t *dst, const KidPtrList &src, int n_src, uint64_t word0,
                               uint64_t n_words, cudaStream_t stream);
cudaError_t kid_launch_lookup(const KidTableView &t, const uint64_t *keys, size_t n, uint32_t *out,
                              cudaStream_t stream);
cudaError_t kid_launch_msca(const KidTreeView &t, const int32_t *x, const int32_t *y, size_t n,
                            int32_t *out, cudaStream_t stream);
