// kid_kernels.cuh - launch interfaces between kid_api.cu and the kernel files.
#pragma once
#include "kid_common.cuh"
#include "kid_table2.cuh"

#define KID_KSIZE 30
#define KID_CLASSIFY_THREADS 256
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

// every kernel launch of this library bumps this (bench.py reports it as gpu_launches)
extern unsigned long long g_kid_kernel_launches;
#define KID_COUNT_LAUNCH() (__atomic_add_fetch(&g_kid_kernel_launches, 1ULL, __ATOMIC_RELAXED))

cudaError_t kid_launch_classify(const KidClassifyParams &p, int sm_count, cudaStream_t stream);  // layout K
cudaError_t kid_launch_classify2(const KidClassifyParams &p, int sm_count, cudaStream_t stream); // layout M

// ---- table build (kid_build.cu) -----------------------------------------------------------------
struct KidBuildStatus {
    unsigned long long n_distinct;  // slots claimed
    unsigned long long n_displaced; // claimed outside the home bucket
    unsigned int range_error;       // some taxa[i] >= n_taxa
    unsigned int overflow;          // some key found no slot within KID_MAX_DISP buckets
};
// slots must be zeroed, owner filled with 0xFFFFFFFF, status zeroed
cudaError_t kid_launch_build(uint64_t *slots, int log2_buckets, uint32_t *owner,
                             const uint64_t *keys, const uint32_t *taxa, size_t n_keys, int n_taxa,
                             KidBuildStatus *status, cudaStream_t stream);

// layout M: owner (uint32 per slot, 3 per sector) must be filled with 0xFFFFFFFF, status zeroed
struct Kid2BuildStatus {
    unsigned long long n_distinct;
    unsigned long long n_displaced; // keys outside their home sector
    unsigned int max_probe;
    unsigned int range_error;
    unsigned int overflow;          // a key needed more than KID2_BUILD_MAX_PROBE sectors
    unsigned int pad;
};
#define KID2_BUILD_MAX_PROBE 4096
cudaError_t kid_launch_build2(uint4 *sectors, int log2_lines, uint32_t *owner, const uint64_t *keys,
                              const uint32_t *taxa, size_t n_keys, int n_taxa, Kid2BuildStatus *status,
                              cudaStream_t stream);
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
cudaError_t kid_launch_seen_or(uint32_t *dst, const KidPtrList &src, int n_src, uint64_t word0,
                               uint64_t n_words, cudaStream_t stream);
cudaError_t kid_launch_lookup(const KidTableView &t, const uint64_t *keys, size_t n, uint32_t *out,
                              cudaStream_t stream);
cudaError_t kid_launch_msca(const KidTreeView &t, const int32_t *x, const int32_t *y, size_t n,
                            int32_t *out, cudaStream_t stream);
