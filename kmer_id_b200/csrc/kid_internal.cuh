// kid_internal.cuh - what the translation units behind the C-ABI (kid_api.cu, kid_ingest.cu) share:
// the objects behind the opaque handles of include/kmer_id.h and the error helpers.
#pragma once
#include "../../include/kmer_id.h"
#include "kid_kernels.cuh"

// records the message kid_last_error() returns for the calling thread; returns `code`
int kid_fail(int code, const char *fmt, ...);

#define KID_CUDA(call)                                                                            \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return kid_fail(e_ == cudaErrorMemoryAllocation ? KID_ENOMEM : KID_ECUDA, "%s: %s (%s:%d)", \
                            #call, cudaGetErrorString(e_), __FILE__, __LINE__);                   \
    } while (0)

struct DeviceGuard { // make `device` current for the call, restore the caller's device after
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int device)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        ok = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

struct HostSlot { // one asynchronous slot: a stream and its device staging buffers
    cudaStream_t stream = nullptr;
    uint8_t *seq = nullptr, *qual = nullptr; // text batches
    uint64_t *off = nullptr;
    uint32_t *words = nullptr;               // packed batch: copied from the host or written by kid_pack_kernel
    uint2 *meta = nullptr;
    int32_t *out_taxon = nullptr;
    uint32_t *out_span = nullptr;
    size_t cap_bytes = 0, cap_reads = 0, cap_words = 0;
    bool has_qual = false;
    uint32_t *dn_codes = nullptr, *dn_boff = nullptr, *dn_flags = nullptr, *dn_inv = nullptr; // dense batches
    size_t dn_cap_codes = 0, dn_cap_reads = 0, dn_cap_inv = 0;
};

struct kid_db {
    int device = 0;
    int n_taxa = 0;
    int layout = KID_LAYOUT_MINIMIZER;
    int log2_sectors = 0;
    int sm_count = 148;
    int max_probe = 0;
    int sub_bits = 2;           // layout M: log2(sectors per minimizer-addressed group)
    int mm = 16;                // layout M: minimizer length (16, or 20 for very large databases)
    unsigned flags = 0;
    uint64_t n_sectors = 0;     // addressable home sectors of 32 bytes (K: 4 slots each, M: 3 entries each)
    uint64_t total_sectors = 0; // n_sectors + slack (clusters run past the last home sector, no wrap)
    uint64_t *slots = nullptr;  // layout K
    uint4 *entries = nullptr;   // layout M
    uint2 *tree = nullptr;
    uint64_t n_distinct = 0, n_displaced = 0;

    uint64_t n_slots() const { return (layout == KID_LAYOUT_KEYHASH ? 4 : KID2_SLOTS_PER_SECTOR) * total_sectors; }
    const void *table_ptr() const { return layout == KID_LAYOUT_KEYHASH ? (const void *)slots : (const void *)entries; }
    KidTableView table_view() const { return KidTableView{ slots, n_sectors - 1, 60 - log2_sectors }; }
    Kid2TableView table_view2() const { return Kid2TableView{ entries, n_sectors - 1, 32 - (log2_sectors - sub_bits), max_probe, sub_bits, mm }; }
    KidTreeView tree_view() const { return KidTreeView{ tree, n_taxa }; }
};

struct kid_sample {
    const kid_db *db = nullptr;
    int *gcount = nullptr, *ucount = nullptr;
    uint32_t *seen = nullptr;      // the bitmap in use (own_seen or a caller-provided buffer)
    uint32_t *own_seen = nullptr;  // what this object allocated
    uint64_t n_words = 0;
    unsigned long long *counters = nullptr;
    cudaEvent_t begin_ev = nullptr;
    HostSlot slot[KID_MAX_SLOTS];
    HostSlot dev; // scratch of kid_classify_device (packed form of a text batch); no stream of its own
    size_t chunk_reads = (size_t)1 << 18;
    uint64_t h2d = 0, d2h = 0;
};

inline KidPackedParams kid_make_packed_params(kid_sample *s, const uint32_t *words, const uint2 *meta, uint32_t bias,
                                              size_t n, int32_t *out_taxon)
{
    KidPackedParams p;
    p.table2 = s->db->table_view2();
    p.tree = s->db->tree_view();
    p.words = words;
    p.meta = meta;
    p.word_bias = bias;
    p.n_reads = n;
    p.out_taxon = out_taxon;
    p.gcount = s->gcount;
    p.seen = s->seen;
    p.counters = s->counters;
    return p;
}
