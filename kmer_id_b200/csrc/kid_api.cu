// kid_api.cu - the C-ABI of include/kmer_id.h: contexts, memory, streams and kernel launches.
// No CPU fallback: every entry point needs a CUDA device and says so when there is none.
#include "../../include/kmer_id.h"
#include "kid_internal.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

unsigned long long g_kid_kernel_launches = 0;

namespace {

thread_local char g_err[512] = "";

// device whose context page-locks host memory for kid_host_alloc: the last one a database was built
// on or kid_device_init was called for (so that a host using GPU 3 does not wake GPU 0 for it)
int g_host_alloc_device = 0;

} // namespace

int kid_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define fail kid_fail

extern "C" {

const char *kid_last_error(void) { return g_err; }

const char *kid_version(void) { return "kmer_id_b200 0.1 (sm_100a)"; }

int kid_device_count(int *n)
{
    if (!n) return fail(KID_EINVAL, "kid_device_count: n is NULL");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        *n = 0;
        return fail(KID_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *n = c;
    return KID_OK;
}

int kid_device_init(int device)
{
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) return fail(KID_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    if (device < 0 || device >= c) return fail(KID_EINVAL, "kid_device_init: device %d of %d", device, c);
    DeviceGuard guard(device);
    // KID_SYNC_BLOCKING=1: host threads sleep in cudaStreamSynchronize instead of spinning (hosts with fewer
    // cores than threads waiting on the GPU); must be set before the context exists, ignored afterwards
    if (getenv("KID_SYNC_BLOCKING")) cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync);
    e = cudaFree(nullptr); // forces the primary context into existence
    if (e != cudaSuccess) return fail(KID_ECUDA, "kid_device_init: %s", cudaGetErrorString(e));
    __atomic_store_n(&g_host_alloc_device, device, __ATOMIC_RELAXED);
    return KID_OK;
}

unsigned long long kid_kernel_launches(void)
{
    return __atomic_load_n(&g_kid_kernel_launches, __ATOMIC_RELAXED);
}

int kid_host_alloc(void **p, size_t bytes)
{
    if (!p) return fail(KID_EINVAL, "kid_host_alloc: p is NULL");
    *p = nullptr;
    DeviceGuard guard(__atomic_load_n(&g_host_alloc_device, __ATOMIC_RELAXED));
    cudaError_t e = cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess)
        return fail(e == cudaErrorMemoryAllocation ? KID_ENOMEM : KID_ECUDA, "cudaHostAlloc(%zu): %s", bytes,
                    cudaGetErrorString(e));
    return KID_OK;
}

void kid_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------------------------------ db
static int build_tree_host(const int32_t *parent, int n_taxa, std::vector<uint2> &node)
{
    // Tree1::get_parent (:146-152): nodes 0 and 1 always report the root, whatever parent[] says
    node.assign((size_t)n_taxa, make_uint2(1u, 0xFFFFFFFFu));
    for (int v = 0; v < n_taxa; v++) {
        int pv = (v != 1 && v > 0) ? parent[v] : 1;
        if (pv < 0 || pv >= n_taxa)
            return fail(KID_ERANGE, "taxonomy: parent[%d] = %d is outside [0,%d)", v, pv, n_taxa);
        node[(size_t)v].x = (uint32_t)pv;
    }
    node[1].y = 0;
    std::vector<int> chain;
    for (int v = 0; v < n_taxa; v++) {
        if (node[(size_t)v].y != 0xFFFFFFFFu) continue;
        chain.clear();
        int w = v;
        while (node[(size_t)w].y == 0xFFFFFFFFu) {
            chain.push_back(w);
            if ((int)chain.size() > n_taxa)
                return fail(KID_ETREE, "taxonomy: node %d never reaches the root (cycle)", v);
            w = (int)node[(size_t)w].x;
        }
        uint32_t d = node[(size_t)w].y;
        for (size_t i = chain.size(); i-- > 0;) node[(size_t)chain[i]].y = ++d;
    }
    return KID_OK;
}

int kid_db_build(const uint64_t *keys, const uint32_t *taxa, size_t n_keys, int keys_on_device,
                 const int32_t *parent, int n_taxa, int device, unsigned flags, int log2_sectors,
                 void *stream_, kid_db **out)
{
    if (!out) return fail(KID_EINVAL, "kid_db_build: out is NULL");
    *out = nullptr;
    const int max_taxa = (flags & KID_DB_LAYOUT_KEYHASH) ? KID_MAX_TAXA : KID2_MAX_TAXA + 1;
    if (n_taxa < 2 || n_taxa > max_taxa)
        return fail(KID_EINVAL, "kid_db_build: n_taxa %d outside [2,%d]", n_taxa, max_taxa);
    if (!parent) return fail(KID_EINVAL, "kid_db_build: parent is NULL");
    if (n_keys && (!keys || !taxa)) return fail(KID_EINVAL, "kid_db_build: keys/taxa NULL");
    if (n_keys >= 0xFFFFFFFFull) return fail(KID_EINVAL, "kid_db_build: more than 2^32-2 probe entries");
    const int layout = (flags & KID_DB_LAYOUT_KEYHASH) ? KID_LAYOUT_KEYHASH : KID_LAYOUT_MINIMIZER;
    const int lo = layout == KID_LAYOUT_KEYHASH ? KID_MIN_LOG2_BUCKETS : KID2_MIN_LOG2_LINES + 2;
    const int hi = layout == KID_LAYOUT_KEYHASH ? KID_MAX_LOG2_BUCKETS : KID2_MAX_LOG2_LINES + 2;
    if (log2_sectors && (log2_sectors < lo || log2_sectors > hi))
        return fail(KID_EINVAL, "kid_db_build: log2_sectors %d outside [%d,%d] for this layout", log2_sectors, lo, hi);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(KID_ECUDA, "kid_db_build: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(KID_EINVAL, "kid_db_build: device %d of %d", device, ndev);

    std::vector<uint2> node;
    int rc = build_tree_host(parent, n_taxa, node);
    if (rc) return rc;

    DeviceGuard guard(device);
    if (!guard.ok) return fail(KID_ECUDA, "cudaSetDevice(%d) failed", device);
    __atomic_store_n(&g_host_alloc_device, device, __ATOMIC_RELAXED);
    cudaStream_t stream = (cudaStream_t)stream_;

    const bool fixed = log2_sectors != 0;
    int B = log2_sectors;
    if (!fixed) {
        // layout K: ~1 key per 4-slot sector.  layout M: <= 0.45 keys per 3-slot sector, so that <1 % of
        // lookups meet a full sector and need a second one
        B = lo;
        const double per_sector = layout == KID_LAYOUT_KEYHASH ? 1.0 : 0.45;
        while (B < hi && (double)((uint64_t)1 << B) * per_sector < (double)n_keys) B++;
        // table + build scratch (owner array, sort buffers): ~44 B per sector + 36 B per key
        const double slots = layout == KID_LAYOUT_KEYHASH ? 4.0 : (double)KID2_SLOTS_PER_SECTOR;
        auto footprint = [&](int b) { return (double)((uint64_t)1 << b) * (32.0 + 4.0 * slots) + 36.0 * (double)n_keys; };
        // layout M: half the load again.  A warp of 32 lookups then rarely has a lane that needs the
        // second, dependent sector load (measured +3 % lookups/s at bact10 scale: 8.6 -> 17 GB of 180).
        // KID_DB_DENSE=1 keeps the denser table.  The size is a function of n_keys alone, so that
        // replicas on several GPUs agree slot for slot (the cross-GPU ucount reduction relies on it) ...
        if (layout == KID_LAYOUT_MINIMIZER && B > lo && B < hi && getenv("KID_DB_DENSE") == nullptr) B++;
        // ... unless the device cannot hold it: then it shrinks, and says so
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            const int want = B;
            while (B > lo && footprint(B) > 0.85 * (double)free_b &&
                   (double)((uint64_t)1 << (B - 1)) * slots * 0.8 > (double)n_keys)
                B--;
            if (B != want && n_keys < 400000000ull) // (beyond that the nominal size exceeds any one GPU)
                fprintf(stderr, "kmer_id_b200: device %d has %.1f GB free: probe table 2^%d sectors instead of 2^%d\n",
                        device, (double)free_b / 1e9, B, want);
        }
    }

    // Beyond 4e8 keys the m = 16 minimizers saturate (kid_table2.cuh): a minimizer then addresses 2 or 4
    // lines instead of one (KID_DB_SUB_BITS=2|3|4 overrides).  KID_DB_MM=20 selects 20-mer minimizers
    // instead (one line per minimizer again; each candidate carries an (order, identity) pair through
    // the sliding minimum, kid_table2.cuh).  Measured at 1.09e9 keys in 2^31 sectors
    // (tools/gpu_runs/gpu_t.sh): m = 16 with two lines per minimizer 19.5 M keys displaced (1.8 %), 110 G
    // lookups/s; m = 20: 4.6 M displaced (0.42 %) but 101 G lookups/s - the 64-bit sliding minimum and the
    // shorter runs (11 windows instead of 15: more DRAM lines per warp) cost more than the second probes
    // they save.  Hence m = 16 stays the default at every size.
    int sub_bits = 2, mm = 16;
    if (layout == KID_LAYOUT_MINIMIZER) {
        if (const char *e = getenv("KID_DB_MM")) {
            const int v = atoi(e);
            if (v == 16 || v == KID_MM20) mm = v;
        }
        const double per_line = (double)n_keys * 4.0 / (double)((uint64_t)1 << B);
        if (mm == 16 && n_keys > 400000000ull) {
            if (per_line > 3.0) sub_bits = 4;
            else if (per_line > 1.5) sub_bits = 3;
        }
        if (const char *e = getenv("KID_DB_SUB_BITS")) {
            const int v = atoi(e);
            if (v >= 2 && v <= 4) sub_bits = v;
        }
        if (sub_bits > B - 2) sub_bits = 2; // tiny tables
    }

    kid_db *db = new (std::nothrow) kid_db;
    if (!db) return fail(KID_ENOMEM, "kid_db_build: host allocation failed");
    db->device = device;
    db->n_taxa = n_taxa;
    db->flags = flags;
    db->layout = layout;
    cudaDeviceGetAttribute(&db->sm_count, cudaDevAttrMultiProcessorCount, device);

    const uint64_t *dkeys = keys;
    const uint32_t *dtaxa = taxa;
    uint64_t *tmp_keys = nullptr;
    uint32_t *tmp_taxa = nullptr, *owner = nullptr;
    void *dstatus = nullptr;
    auto cleanup_tmp = [&]() {
        cudaFree(tmp_keys); cudaFree(tmp_taxa); cudaFree(owner); cudaFree(dstatus);
        tmp_keys = nullptr; tmp_taxa = nullptr; owner = nullptr; dstatus = nullptr;
    };
    auto bail = [&](int code) { cleanup_tmp(); kid_db_free(db); return code; };
#define KID_CUDA_B(call)                                                                          \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return bail(fail(e_ == cudaErrorMemoryAllocation ? KID_ENOMEM : KID_ECUDA,            \
                             "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); \
    } while (0)

    KID_CUDA_B(cudaMalloc(&db->tree, sizeof(uint2) * (size_t)n_taxa));
    KID_CUDA_B(cudaMemcpyAsync(db->tree, node.data(), sizeof(uint2) * (size_t)n_taxa,
                               cudaMemcpyHostToDevice, stream));
    if (n_keys && !keys_on_device) {
        KID_CUDA_B(cudaMalloc(&tmp_keys, sizeof(uint64_t) * n_keys));
        KID_CUDA_B(cudaMalloc(&tmp_taxa, sizeof(uint32_t) * n_keys));
        KID_CUDA_B(cudaMemcpyAsync(tmp_keys, keys, sizeof(uint64_t) * n_keys, cudaMemcpyHostToDevice, stream));
        KID_CUDA_B(cudaMemcpyAsync(tmp_taxa, taxa, sizeof(uint32_t) * n_keys, cudaMemcpyHostToDevice, stream));
        dkeys = tmp_keys;
        dtaxa = tmp_taxa;
    }
    KID_CUDA_B(cudaMalloc(&dstatus, 64));

    for (;;) {
        const uint64_t n_sectors = (uint64_t)1 << B;
        KidSortedBuildParams bp;
        bp.layout = layout;
        bp.slots_per_sector = layout == KID_LAYOUT_KEYHASH ? 4 : KID2_SLOTS_PER_SECTOR;
        bp.n_sectors = n_sectors;
        bp.slack_sectors = layout == KID_LAYOUT_KEYHASH ? KID1_SLACK_SECTORS : KID2_SLACK_SECTORS;
        bp.sub_bits = sub_bits;
        bp.mm = mm;
        bp.line_shift = 32 - (B - sub_bits);
        bp.rem_bits = 60 - B;
        bp.n_taxa = (uint32_t)n_taxa;
        bp.max_taxon = layout == KID_LAYOUT_KEYHASH ? (uint32_t)KID_MAX_TAXA : (uint32_t)KID2_MAX_TAXA;
        bp.max_disp = layout == KID_LAYOUT_KEYHASH ? (uint32_t)KID_MAX_DISP : (uint32_t)(KID2_SLACK_SECTORS - 1);
        const uint64_t total_sectors = n_sectors + bp.slack_sectors;
        const size_t n_slots = (size_t)(total_sectors * (uint64_t)bp.slots_per_sector);
        Kid2BuildStatus st;
        KID_CUDA_B(cudaMalloc(&owner, n_slots * sizeof(uint32_t)));
        KID_CUDA_B(kid_build_owner_sorted(dkeys, dtaxa, n_keys, bp, owner, static_cast<Kid2BuildStatus *>(dstatus),
                                          &st, stream));
        if (st.range_error)
            return bail(fail(KID_ERANGE, "probe table: a probe names a taxon >= n_taxa (%d); the "
                                         "reference indexes gcount[] out of bounds here", n_taxa));
        if (!st.overflow) {
            if (layout == KID_LAYOUT_KEYHASH) {
                KID_CUDA_B(cudaMalloc(&db->slots, n_slots * sizeof(uint64_t)));
                KID_CUDA_B(kid_launch_pack1(db->slots, n_slots, bp.rem_bits, owner, dkeys, dtaxa, stream));
            } else {
                KID_CUDA_B(cudaMalloc(&db->entries, (size_t)(2 * total_sectors) * sizeof(uint4)));
                KID_CUDA_B(kid_launch_pack2(db->entries, (size_t)total_sectors, owner, dkeys, dtaxa, stream));
            }
            KID_CUDA_B(cudaStreamSynchronize(stream));
            db->log2_sectors = B;
            db->sub_bits = sub_bits;
            db->mm = mm;
            db->n_sectors = n_sectors;
            db->total_sectors = total_sectors;
            db->n_distinct = st.n_distinct;
            db->n_displaced = st.n_displaced;
            db->max_probe = (int)st.max_probe;
            cudaFree(owner);
            owner = nullptr;
            break;
        }
        cudaFree(owner);
        owner = nullptr;
        if (fixed || B == hi)
            return bail(fail(KID_EFULL, "probe table: 2^%d sectors cannot place every key", B));
        B++;
    }
    cleanup_tmp();
    *out = db;
    return KID_OK;
#undef KID_CUDA_B
}

void kid_db_free(kid_db *db)
{
    if (!db) return;
    DeviceGuard guard(db->device);
    cudaFree(db->slots);
    cudaFree(db->entries);
    cudaFree(db->tree);
    delete db;
}

int kid_db_n_taxa(const kid_db *db) { return db ? db->n_taxa : 0; }
int kid_db_device(const kid_db *db) { return db ? db->device : -1; }

int kid_db_stats(const kid_db *db, uint64_t *n_distinct, uint64_t *n_buckets, uint64_t *table_bytes,
                 uint64_t *n_displaced)
{
    if (!db) return fail(KID_EINVAL, "kid_db_stats: db is NULL");
    if (n_distinct) *n_distinct = db->n_distinct;
    if (n_buckets) *n_buckets = db->n_sectors;
    if (table_bytes) *table_bytes = db->total_sectors * 32;
    if (n_displaced) *n_displaced = db->n_displaced;
    return KID_OK;
}

int kid_db_table_device(const kid_db *db, void **table, uint64_t *n_buckets)
{
    if (!db || !table) return fail(KID_EINVAL, "kid_db_table_device: NULL argument");
    *table = const_cast<void *>(db->table_ptr());
    if (n_buckets) *n_buckets = db->n_sectors;
    return KID_OK;
}

int kid_db_lookup(const kid_db *db, const uint64_t *keys, size_t n, uint32_t *taxa_out)
{
    if (!db || (n && (!keys || !taxa_out))) return fail(KID_EINVAL, "kid_db_lookup: NULL argument");
    if (n == 0) return KID_OK;
    DeviceGuard guard(db->device);
    uint64_t *dk = nullptr;
    uint32_t *dt = nullptr;
    KID_CUDA(cudaMalloc(&dk, n * sizeof(uint64_t)));
    cudaError_t e = cudaMalloc(&dt, n * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemcpy(dk, keys, n * sizeof(uint64_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = db->layout == KID_LAYOUT_KEYHASH ? kid_launch_lookup(db->table_view(), dk, n, dt, nullptr)
                                             : kid_launch_lookup2(db->table_view2(), dk, n, dt, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(taxa_out, dt, n * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    cudaFree(dk);
    cudaFree(dt);
    if (e != cudaSuccess) return fail(KID_ECUDA, "kid_db_lookup: %s", cudaGetErrorString(e));
    return KID_OK;
}

int kid_db_msca(const kid_db *db, const int32_t *x, const int32_t *y, size_t n, int32_t *out)
{
    if (!db || (n && (!x || !y || !out))) return fail(KID_EINVAL, "kid_db_msca: NULL argument");
    for (size_t i = 0; i < n; i++)
        if (x[i] < 0 || x[i] >= db->n_taxa || y[i] < 0 || y[i] >= db->n_taxa)
            return fail(KID_ERANGE, "kid_db_msca: pair %zu (%d,%d) outside [0,%d)", i, x[i], y[i], db->n_taxa);
    if (n == 0) return KID_OK;
    DeviceGuard guard(db->device);
    int32_t *d = nullptr;
    KID_CUDA(cudaMalloc(&d, 3 * n * sizeof(int32_t)));
    cudaError_t e = cudaMemcpy(d, x, n * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + n, y, n * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = kid_launch_msca(db->tree_view(), d, d + n, n, d + 2 * n, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(out, d + 2 * n, n * sizeof(int32_t), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(KID_ECUDA, "kid_db_msca: %s", cudaGetErrorString(e));
    return KID_OK;
}

// -------------------------------------------------------------------------------------- sample
int kid_sample_create(const kid_db *db, kid_sample **out)
{
    if (!db || !out) return fail(KID_EINVAL, "kid_sample_create: NULL argument");
    *out = nullptr;
    DeviceGuard guard(db->device);
    kid_sample *s = new (std::nothrow) kid_sample;
    if (!s) return fail(KID_ENOMEM, "kid_sample_create: host allocation failed");
    s->db = db;
    const uint64_t n_slots = db->n_slots();
    s->n_words = ((n_slots / 32) + 1023) / 1024 * 1024;
    cudaError_t e = cudaMalloc(&s->gcount, sizeof(int) * (size_t)db->n_taxa);
    if (e == cudaSuccess) e = cudaMalloc(&s->ucount, sizeof(int) * (size_t)db->n_taxa);
    if (e == cudaSuccess) e = cudaMalloc(&s->own_seen, sizeof(uint32_t) * s->n_words);
    s->seen = s->own_seen;
    if (e == cudaSuccess) e = cudaMalloc(&s->counters, 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->begin_ev, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        kid_sample_free(s);
        return fail(e == cudaErrorMemoryAllocation ? KID_ENOMEM : KID_ECUDA, "kid_sample_create: %s",
                    cudaGetErrorString(e));
    }
    int rc = kid_sample_begin(s, nullptr);
    if (rc) { kid_sample_free(s); return rc; }
    KID_CUDA(cudaStreamSynchronize(nullptr));
    *out = s;
    return KID_OK;
}

void kid_sample_free(kid_sample *s)
{
    if (!s) return;
    DeviceGuard guard(s->db->device);
    for (HostSlot &h : s->slot) {
        if (h.stream) { cudaStreamSynchronize(h.stream); cudaStreamDestroy(h.stream); }
        cudaFree(h.seq); cudaFree(h.qual); cudaFree(h.off); cudaFree(h.out_taxon); cudaFree(h.out_span);
        cudaFree(h.words); cudaFree(h.meta);
        cudaFree(h.dn_codes); cudaFree(h.dn_boff); cudaFree(h.dn_flags); cudaFree(h.dn_inv);
    }
    cudaFree(s->dev.words); cudaFree(s->dev.meta);
    if (s->begin_ev) cudaEventDestroy(s->begin_ev);
    cudaFree(s->gcount); cudaFree(s->ucount); cudaFree(s->own_seen); cudaFree(s->counters);
    delete s;
}

const kid_db *kid_sample_db(const kid_sample *s) { return s ? s->db : nullptr; }

int kid_sample_begin(kid_sample *s, void *stream_)
{
    if (!s) return fail(KID_EINVAL, "kid_sample_begin: s is NULL");
    DeviceGuard guard(s->db->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    KID_CUDA(cudaMemsetAsync(s->gcount, 0, sizeof(int) * (size_t)s->db->n_taxa, stream));
    KID_CUDA(cudaMemsetAsync(s->seen, 0, sizeof(uint32_t) * s->n_words, stream));
    KID_CUDA(cudaMemsetAsync(s->counters, 0, 2 * sizeof(unsigned long long), stream));
    KID_CUDA(cudaEventRecord(s->begin_ev, stream));
    s->h2d = s->d2h = 0;
    return KID_OK;
}

static cudaError_t launch_classify(const kid_db *db, const KidClassifyParams &p, cudaStream_t stream)
{
    return kid_launch_classify(p, db->sm_count, stream); // layout K only
}

static KidClassifyParams make_params(kid_sample *s, const uint8_t *seq, const uint8_t *qual,
                                     const uint64_t *off, uint64_t bias, size_t n, int32_t *out_taxon,
                                     uint32_t *out_span)
{
    KidClassifyParams p;
    p.table = s->db->table_view();
    p.table2 = s->db->table_view2();
    p.tree = s->db->tree_view();
    p.seq = seq;
    p.qual = qual;
    p.off = off;
    p.off_bias = bias;
    p.n_reads = n;
    p.out_taxon = out_taxon;
    p.out_span = out_span;
    p.gcount = s->gcount;
    p.seen = s->seen;
    p.counters = s->counters;
    p.accept_u = (s->db->flags & KID_DB_ACCEPT_U) != 0;
    return p;
}

// layout K (the key-hashed bake-off table) keeps its own fused text kernel; layout M packs, then scans
static bool use_fused_text_kernel(const kid_db *db) { return db->layout == KID_LAYOUT_KEYHASH; }

static int reserve_packed(HostSlot &h, size_t words, size_t reads)
{
    if (words + 16 > h.cap_words) {
        if (h.stream) KID_CUDA(cudaStreamSynchronize(h.stream));
        cudaFree(h.words);
        h.words = nullptr;
        h.cap_words = 0;
        const size_t cap = words + words / 4 + 16;
        KID_CUDA(cudaMalloc(&h.words, sizeof(uint32_t) * cap));
        h.cap_words = cap;
    }
    if (reads > h.cap_reads || !h.meta) {
        if (h.stream) KID_CUDA(cudaStreamSynchronize(h.stream));
        cudaFree(h.off); cudaFree(h.out_taxon); cudaFree(h.out_span); cudaFree(h.meta);
        h.off = nullptr; h.out_taxon = nullptr; h.out_span = nullptr; h.meta = nullptr;
        h.cap_reads = 0;
        const size_t cap = reads + reads / 4 + 16;
        KID_CUDA(cudaMalloc(&h.off, sizeof(uint64_t) * (cap + 1)));
        KID_CUDA(cudaMalloc(&h.meta, sizeof(uint2) * (cap + 1)));
        KID_CUDA(cudaMalloc(&h.out_taxon, sizeof(int32_t) * cap));
        KID_CUDA(cudaMalloc(&h.out_span, sizeof(uint32_t) * 2 * cap));
        h.cap_reads = cap;
    }
    return KID_OK;
}

static int reserve_text(HostSlot &h, size_t bytes, bool want_qual)
{
    if (bytes + 64 > h.cap_bytes || (want_qual && !h.has_qual)) {
        if (h.stream) KID_CUDA(cudaStreamSynchronize(h.stream));
        cudaFree(h.seq); cudaFree(h.qual);
        h.seq = h.qual = nullptr;
        size_t cap = bytes + bytes / 4 + 64;
        if (cap < h.cap_bytes) cap = h.cap_bytes;
        h.cap_bytes = 0;
        h.has_qual = false;
        KID_CUDA(cudaMalloc(&h.seq, cap));
        if (want_qual) KID_CUDA(cudaMalloc(&h.qual, cap));
        h.has_qual = want_qual;
        h.cap_bytes = cap;
    }
    return KID_OK;
}

// text on the device -> packed form in h.words/h.meta -> k-mer scan, all on `stream`
static int pack_and_scan(kid_sample *s, HostSlot &h, const uint8_t *seq, const uint8_t *qual, const uint64_t *off,
                         uint64_t bias, uint64_t bytes, size_t n, int32_t *out_taxon, uint32_t *out_span,
                         cudaStream_t stream)
{
    const uint64_t bound = kid_pack_word_index(bytes, n) + 2;
    if (bound >= 0x80000000ull)
        return fail(KID_ERANGE, "a text batch of %llu bases in %zu reads needs more than 2^31 packed words; split it",
                    (unsigned long long)bytes, n);
    int rc = reserve_packed(h, (size_t)bound, n);
    if (rc) return rc;
    KidPackParams pp;
    pp.seq = seq;
    pp.qual = qual;
    pp.off = off;
    pp.off_bias = bias;
    pp.n_reads = n;
    pp.words = h.words;
    pp.meta = h.meta;
    pp.out_span = out_span;
    pp.accept_u = (s->db->flags & KID_DB_ACCEPT_U) != 0;
    KID_CUDA(kid_launch_pack(pp, s->db->sm_count, stream));
    KidPackedParams p = kid_make_packed_params(s, h.words, h.meta, 0, n, out_taxon);
    KID_CUDA(kid_launch_classify3(p, s->db->sm_count, stream));
    return KID_OK;
}

int kid_classify_device(kid_sample *s, const uint8_t *seq, const uint8_t *qual, const uint64_t *off,
                        size_t n_reads, int32_t *out_taxon, uint32_t *out_span, void *stream_)
{
    if (!s) return fail(KID_EINVAL, "kid_classify_device: s is NULL");
    if (n_reads == 0) return KID_OK;
    if (!seq || !off) return fail(KID_EINVAL, "kid_classify_device: seq/off is NULL");
    if (reinterpret_cast<uintptr_t>(seq) & 15)
        return fail(KID_EINVAL, "kid_classify_device: seq must be 16-byte aligned (the kernels load aligned 128-bit words)");
    DeviceGuard guard(s->db->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (use_fused_text_kernel(s->db)) {
        KidClassifyParams p = make_params(s, seq, qual, off, 0, n_reads, out_taxon, out_span);
        KID_CUDA(launch_classify(s->db, p, stream));
        return KID_OK;
    }
    // the packed form's size depends on the batch's bases: two offsets come back to the host
    uint64_t ends[2] = { 0, 0 };
    KID_CUDA(cudaMemcpyAsync(&ends[0], off, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
    KID_CUDA(cudaMemcpyAsync(&ends[1], off + n_reads, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
    KID_CUDA(cudaStreamSynchronize(stream));
    if (ends[0] != 0) return fail(KID_EINVAL, "kid_classify_device: off[0] must be 0");
    return pack_and_scan(s, s->dev, seq, qual, off, 0, ends[1], n_reads, out_taxon, out_span, stream);
}

int kid_pack_device(const kid_db *db, const uint8_t *seq, const uint8_t *qual, const uint64_t *off,
                    uint64_t total_bases, size_t n_reads, uint32_t *words, size_t words_cap, uint32_t *meta,
                    uint32_t *out_span, void *stream)
{
    if (!db || !meta) return fail(KID_EINVAL, "kid_pack_device: NULL argument");
    if (n_reads && (!seq || !off || !words)) return fail(KID_EINVAL, "kid_pack_device: seq/off/words is NULL");
    if ((reinterpret_cast<uintptr_t>(seq) & 15) || (reinterpret_cast<uintptr_t>(meta) & 7))
        return fail(KID_EINVAL, "kid_pack_device: seq must be 16-byte aligned, meta 8-byte aligned");
    const uint64_t bound = kid_pack_word_index(total_bases, n_reads) + 2;
    if (bound >= 0x80000000ull || bound > words_cap)
        return fail(KID_ERANGE, "kid_pack_device: %llu bases in %zu reads need %llu words (have %zu, limit 2^31)",
                    (unsigned long long)total_bases, n_reads, (unsigned long long)bound, words_cap);
    DeviceGuard guard(db->device);
    KidPackParams pp;
    pp.seq = seq;
    pp.qual = qual;
    pp.off = off;
    pp.off_bias = 0;
    pp.n_reads = n_reads;
    pp.words = words;
    pp.meta = reinterpret_cast<uint2 *>(meta);
    pp.out_span = out_span;
    pp.accept_u = (db->flags & KID_DB_ACCEPT_U) != 0;
    KID_CUDA(kid_launch_pack(pp, db->sm_count, (cudaStream_t)stream));
    return KID_OK;
}

int kid_classify_packed_device(kid_sample *s, const uint32_t *words, const uint32_t *meta, size_t n_reads,
                               int32_t *out_taxon, void *stream)
{
    if (!s) return fail(KID_EINVAL, "kid_classify_packed_device: s is NULL");
    if (n_reads == 0) return KID_OK;
    if (!words || !meta) return fail(KID_EINVAL, "kid_classify_packed_device: words/meta is NULL");
    if (reinterpret_cast<uintptr_t>(meta) & 7) return fail(KID_EINVAL, "kid_classify_packed_device: meta must be 8-byte aligned");
    if (s->db->layout != KID_LAYOUT_MINIMIZER)
        return fail(KID_EINVAL, "packed batches need the default table layout (not KID_DB_LAYOUT_KEYHASH)");
    DeviceGuard guard(s->db->device);
    KidPackedParams p = kid_make_packed_params(s, words, reinterpret_cast<const uint2 *>(meta), 0, n_reads, out_taxon);
    KID_CUDA(kid_launch_classify3(p, s->db->sm_count, (cudaStream_t)stream));
    return KID_OK;
}

int kid_sample_set_chunk_reads(kid_sample *s, size_t chunk_reads)
{
    if (!s || chunk_reads == 0) return fail(KID_EINVAL, "kid_sample_set_chunk_reads: bad argument");
    s->chunk_reads = chunk_reads;
    return KID_OK;
}

int kid_sample_transfer_bytes(const kid_sample *s, uint64_t *h2d, uint64_t *d2h)
{
    if (!s) return fail(KID_EINVAL, "kid_sample_transfer_bytes: s is NULL");
    if (h2d) *h2d = s->h2d;
    if (d2h) *d2h = s->d2h;
    return KID_OK;
}

static int slot_open(kid_sample *s, HostSlot &h)
{
    if (!h.stream) KID_CUDA(cudaStreamCreateWithFlags(&h.stream, cudaStreamNonBlocking));
    // submissions must not overtake kid_sample_begin's memsets
    KID_CUDA(cudaStreamWaitEvent(h.stream, s->begin_ev, 0));
    return KID_OK;
}

// reads [r0, r0+n) of a host text batch: H2D, pack, scan, D2H on the slot's stream
static int submit_text(kid_sample *s, HostSlot &h, const uint8_t *seq, const uint8_t *qual, const uint64_t *off,
                       size_t r0, size_t n, int32_t *out_taxon, uint32_t *out_span)
{
    const uint64_t b0 = off[r0], bytes = off[r0 + n] - b0;
    int rc = slot_open(s, h);
    if (!rc) rc = reserve_text(h, (size_t)bytes, qual != nullptr);
    if (!rc) rc = reserve_packed(h, 0, n);
    if (rc) return rc;
    KID_CUDA(cudaMemcpyAsync(h.seq, seq + b0, bytes, cudaMemcpyHostToDevice, h.stream));
    if (qual) KID_CUDA(cudaMemcpyAsync(h.qual, qual + b0, bytes, cudaMemcpyHostToDevice, h.stream));
    KID_CUDA(cudaMemcpyAsync(h.off, off + r0, sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, h.stream));
    __atomic_fetch_add(&s->h2d, bytes * (qual ? 2 : 1) + sizeof(uint64_t) * (n + 1), __ATOMIC_RELAXED); // (two threads may feed one sample)
    if (use_fused_text_kernel(s->db)) {
        KidClassifyParams p = make_params(s, h.seq, qual ? h.qual : nullptr, h.off, b0, n,
                                          out_taxon ? h.out_taxon : nullptr, out_span ? h.out_span : nullptr);
        KID_CUDA(launch_classify(s->db, p, h.stream));
    } else {
        rc = pack_and_scan(s, h, h.seq, qual ? h.qual : nullptr, h.off, b0, bytes, n, out_taxon ? h.out_taxon : nullptr,
                           out_span ? h.out_span : nullptr, h.stream);
        if (rc) return rc;
    }
    if (out_taxon) {
        KID_CUDA(cudaMemcpyAsync(out_taxon + r0, h.out_taxon, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h.stream));
        __atomic_fetch_add(&s->d2h, sizeof(int32_t) * n, __ATOMIC_RELAXED);
    }
    if (out_span) {
        KID_CUDA(cudaMemcpyAsync(out_span + 2 * r0, h.out_span, sizeof(uint32_t) * 2 * n, cudaMemcpyDeviceToHost, h.stream));
        __atomic_fetch_add(&s->d2h, sizeof(uint32_t) * 2 * n, __ATOMIC_RELAXED);
    }
    return KID_OK;
}

// reads [r0, r0+n) of a host packed batch
static int submit_packed(kid_sample *s, HostSlot &h, const uint32_t *words, uint32_t word0, const uint32_t *meta,
                         size_t r0, size_t n, int32_t *out_taxon)
{
    const uint32_t wa = meta[2 * r0] & ~KID_PK_INVALID, wb = meta[2 * (r0 + n)] & ~KID_PK_INVALID;
    if (wa < word0 || wb < wa) return fail(KID_EINVAL, "packed batch: word indices must be non-decreasing and >= word0");
    const size_t nw = wb - wa;
    int rc = slot_open(s, h);
    if (!rc) rc = reserve_packed(h, nw, n);
    if (rc) return rc;
    if (nw) KID_CUDA(cudaMemcpyAsync(h.words, words + (wa - word0), sizeof(uint32_t) * nw, cudaMemcpyHostToDevice, h.stream));
    KID_CUDA(cudaMemcpyAsync(h.meta, meta + 2 * r0, sizeof(uint2) * (n + 1), cudaMemcpyHostToDevice, h.stream));
    __atomic_fetch_add(&s->h2d, sizeof(uint32_t) * nw + sizeof(uint2) * (n + 1), __ATOMIC_RELAXED);
    KidPackedParams p = kid_make_packed_params(s, h.words, h.meta, wa, n, out_taxon ? h.out_taxon : nullptr);
    KID_CUDA(kid_launch_classify3(p, s->db->sm_count, h.stream));
    if (out_taxon) {
        KID_CUDA(cudaMemcpyAsync(out_taxon + r0, h.out_taxon, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h.stream));
        __atomic_fetch_add(&s->d2h, sizeof(int32_t) * n, __ATOMIC_RELAXED);
    }
    return KID_OK;
}

// kid_classify_host / kid_classify_packed_host rotate their chunks over this many slots: while one
// chunk's kernel runs, the copies of the next ones keep the DMA engine busy
static const int kHostSlots = getenv("KID_HOST_SLOTS") ? std::max(1, std::min(KID_MAX_SLOTS, atoi(getenv("KID_HOST_SLOTS")))) : 4; // measured: 2 -> 447, 3 -> 472, 4 -> 485 M pairs/s

static int reserve_dense(HostSlot &h, size_t code_words, size_t reads, size_t n_inv)
{
    auto grow = [&](uint32_t *&p, size_t &cap, size_t need) -> int {
        if (need <= cap && p) return KID_OK;
        if (h.stream) KID_CUDA(cudaStreamSynchronize(h.stream));
        cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t c = need + need / 4 + 64;
        KID_CUDA(cudaMalloc(&p, sizeof(uint32_t) * c));
        cap = c;
        return KID_OK;
    };
    int rc = grow(h.dn_codes, h.dn_cap_codes, code_words + 8);
    if (!rc) rc = grow(h.dn_inv, h.dn_cap_inv, n_inv + 1);
    if (!rc && (reads + 1 > h.dn_cap_reads || !h.dn_boff)) {
        if (h.stream) KID_CUDA(cudaStreamSynchronize(h.stream));
        cudaFree(h.dn_boff); cudaFree(h.dn_flags);
        h.dn_boff = h.dn_flags = nullptr;
        h.dn_cap_reads = 0;
        const size_t c = reads + reads / 4 + 64;
        KID_CUDA(cudaMalloc(&h.dn_boff, sizeof(uint32_t) * (c + 1)));
        KID_CUDA(cudaMalloc(&h.dn_flags, sizeof(uint32_t) * (c / 32 + 2)));
        h.dn_cap_reads = c;
    }
    return rc;
}

// reads [r0, r0+n) of a host dense batch (r0 a multiple of 32): H2D, expand, scan, D2H on the slot's stream.
// codes[0] holds the stream from position 16*(boff[0]/16) on; inv lists the batch's non-ACGT positions.
static int submit_dense(kid_sample *s, HostSlot &h, const uint32_t *codes, const uint32_t *boff, const uint32_t *flagbits,
                        const uint32_t *inv, size_t n_inv, size_t r0, size_t n, int32_t *out_taxon)
{
    const uint32_t origin = (boff[0] >> 4) << 4;          // stream position of codes[0]
    const uint32_t ba = boff[r0], bb = boff[r0 + n];      // this chunk's bases
    if (bb < ba || ba < origin) return fail(KID_EINVAL, "dense batch: base offsets must not decrease");
    const uint32_t bias = (ba >> 4) << 4;
    const size_t w0 = (bias - origin) >> 4, nw = ((size_t)(bb - bias) + 15) >> 4;
    // this chunk's slice of the position list
    const uint32_t *ia = inv, *ib = inv;
    if (n_inv) {
        ia = std::lower_bound(inv, inv + n_inv, ba);
        ib = std::lower_bound(ia, inv + n_inv, bb);
    }
    const size_t ni = (size_t)(ib - ia);
    const uint64_t bound = kid_pack_word_index((uint64_t)(bb - bias), n) + 2;
    if (bound >= 0x80000000ull) return fail(KID_ERANGE, "dense batch: chunk too large");
    int rc = slot_open(s, h);
    if (!rc) rc = reserve_dense(h, nw, n, ni);
    if (!rc) rc = reserve_packed(h, (size_t)bound, n);
    if (rc) return rc;
    if (nw) KID_CUDA(cudaMemcpyAsync(h.dn_codes, codes + w0, sizeof(uint32_t) * nw, cudaMemcpyHostToDevice, h.stream));
    KID_CUDA(cudaMemcpyAsync(h.dn_boff, boff + r0, sizeof(uint32_t) * (n + 1), cudaMemcpyHostToDevice, h.stream));
    KID_CUDA(cudaMemcpyAsync(h.dn_flags, flagbits + r0 / 32, sizeof(uint32_t) * ((n + 31) / 32), cudaMemcpyHostToDevice, h.stream));
    if (ni) KID_CUDA(cudaMemcpyAsync(h.dn_inv, ia, sizeof(uint32_t) * ni, cudaMemcpyHostToDevice, h.stream));
    __atomic_fetch_add(&s->h2d, sizeof(uint32_t) * (nw + n + 1 + (n + 31) / 32 + ni), __ATOMIC_RELAXED);
    KidExpandParams ep;
    ep.codes = h.dn_codes;
    ep.boff = h.dn_boff;
    ep.bias = bias;
    ep.flagbits = h.dn_flags;
    ep.inv = h.dn_inv;
    ep.n_inv = (uint32_t)ni;
    ep.n_reads = n;
    ep.words = h.words;
    ep.meta = h.meta;
    KID_CUDA(kid_launch_expand(ep, s->db->sm_count, h.stream));
    KidPackedParams p = kid_make_packed_params(s, h.words, h.meta, 0, n, out_taxon ? h.out_taxon : nullptr);
    KID_CUDA(kid_launch_classify3(p, s->db->sm_count, h.stream));
    if (out_taxon) {
        KID_CUDA(cudaMemcpyAsync(out_taxon + r0, h.out_taxon, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h.stream));
        __atomic_fetch_add(&s->d2h, sizeof(int32_t) * n, __ATOMIC_RELAXED);
    }
    return KID_OK;
}

static int sync_slots(kid_sample *s)
{
    for (HostSlot &h : s->slot)
        if (h.stream) KID_CUDA(cudaStreamSynchronize(h.stream));
    return KID_OK;
}

int kid_classify_host(kid_sample *s, const uint8_t *seq, const uint8_t *qual, const uint64_t *off,
                      size_t n_reads, int32_t *out_taxon, uint32_t *out_span)
{
    if (!s) return fail(KID_EINVAL, "kid_classify_host: s is NULL");
    if (n_reads == 0) return KID_OK;
    if (!seq || !off) return fail(KID_EINVAL, "kid_classify_host: seq/off is NULL");
    DeviceGuard guard(s->db->device);
    int k = 0;
    for (size_t r0 = 0; r0 < n_reads; r0 += s->chunk_reads, k = (k + 1) % kHostSlots) {
        const size_t n = (n_reads - r0 < s->chunk_reads) ? n_reads - r0 : s->chunk_reads;
        // the previous chunk on this slot drains before its buffers are overwritten (stream order)
        int rc = submit_text(s, s->slot[k], seq, qual, off, r0, n, out_taxon, out_span);
        if (rc) { sync_slots(s); return rc; }
    }
    return sync_slots(s);
}

int kid_classify_packed_host(kid_sample *s, const uint32_t *words, uint32_t word0, const uint32_t *meta,
                             size_t n_reads, int32_t *out_taxon)
{
    if (!s) return fail(KID_EINVAL, "kid_classify_packed_host: s is NULL");
    if (n_reads == 0) return KID_OK;
    if (!words || !meta) return fail(KID_EINVAL, "kid_classify_packed_host: words/meta is NULL");
    if (s->db->layout != KID_LAYOUT_MINIMIZER)
        return fail(KID_EINVAL, "packed batches need the default table layout (not KID_DB_LAYOUT_KEYHASH)");
    DeviceGuard guard(s->db->device);
    int k = 0;
    for (size_t r0 = 0; r0 < n_reads; r0 += s->chunk_reads, k = (k + 1) % kHostSlots) {
        const size_t n = (n_reads - r0 < s->chunk_reads) ? n_reads - r0 : s->chunk_reads;
        int rc = submit_packed(s, s->slot[k], words, word0, meta, r0, n, out_taxon);
        if (rc) { sync_slots(s); return rc; }
    }
    return sync_slots(s);
}

int kid_classify_dense_host(kid_sample *s, const uint32_t *codes, const uint32_t *boff, const uint32_t *flagbits,
                            const uint32_t *inv, size_t n_inv, size_t n_reads, int32_t *out_taxon)
{
    if (!s) return fail(KID_EINVAL, "kid_classify_dense_host: s is NULL");
    if (n_reads == 0) return KID_OK;
    if (!codes || !boff || !flagbits || (n_inv && !inv)) return fail(KID_EINVAL, "kid_classify_dense_host: NULL argument");
    if (s->db->layout != KID_LAYOUT_MINIMIZER)
        return fail(KID_EINVAL, "dense batches need the default table layout (not KID_DB_LAYOUT_KEYHASH)");
    DeviceGuard guard(s->db->device);
    const size_t chunk = std::max<size_t>(32, s->chunk_reads & ~(size_t)31); // flag words must not straddle chunks
    int k = 0;
    for (size_t r0 = 0; r0 < n_reads; r0 += chunk, k = (k + 1) % kHostSlots) {
        const size_t n = (n_reads - r0 < chunk) ? n_reads - r0 : chunk;
        int rc = submit_dense(s, s->slot[k], codes, boff, flagbits, inv, n_inv, r0, n, out_taxon);
        if (rc) { sync_slots(s); return rc; }
    }
    return sync_slots(s);
}

int kid_classify_dense_async(kid_sample *s, int slot, const uint32_t *codes, const uint32_t *boff, const uint32_t *flagbits,
                             const uint32_t *inv, size_t n_inv, size_t n_reads, int32_t *out_taxon)
{
    if (!s || slot < 0 || slot >= KID_MAX_SLOTS) return fail(KID_EINVAL, "kid_classify_dense_async: bad sample or slot");
    if (n_reads == 0) return KID_OK;
    if (!codes || !boff || !flagbits || (n_inv && !inv)) return fail(KID_EINVAL, "kid_classify_dense_async: NULL argument");
    if (s->db->layout != KID_LAYOUT_MINIMIZER)
        return fail(KID_EINVAL, "dense batches need the default table layout (not KID_DB_LAYOUT_KEYHASH)");
    DeviceGuard guard(s->db->device);
    return submit_dense(s, s->slot[slot], codes, boff, flagbits, inv, n_inv, 0, n_reads, out_taxon);
}

int kid_classify_async(kid_sample *s, int slot, const uint8_t *seq, const uint8_t *qual, const uint64_t *off,
                       size_t n_reads, int32_t *out_taxon, uint32_t *out_span)
{
    if (!s || slot < 0 || slot >= KID_MAX_SLOTS) return fail(KID_EINVAL, "kid_classify_async: bad sample or slot");
    if (n_reads == 0) return KID_OK;
    if (!seq || !off) return fail(KID_EINVAL, "kid_classify_async: seq/off is NULL");
    DeviceGuard guard(s->db->device);
    return submit_text(s, s->slot[slot], seq, qual, off, 0, n_reads, out_taxon, out_span);
}

int kid_classify_packed_async(kid_sample *s, int slot, const uint32_t *words, uint32_t word0, const uint32_t *meta,
                              size_t n_reads, int32_t *out_taxon)
{
    if (!s || slot < 0 || slot >= KID_MAX_SLOTS) return fail(KID_EINVAL, "kid_classify_packed_async: bad sample or slot");
    if (n_reads == 0) return KID_OK;
    if (!words || !meta) return fail(KID_EINVAL, "kid_classify_packed_async: words/meta is NULL");
    if (s->db->layout != KID_LAYOUT_MINIMIZER)
        return fail(KID_EINVAL, "packed batches need the default table layout (not KID_DB_LAYOUT_KEYHASH)");
    DeviceGuard guard(s->db->device);
    return submit_packed(s, s->slot[slot], words, word0, meta, 0, n_reads, out_taxon);
}

int kid_wait(kid_sample *s, int slot)
{
    if (!s || slot < 0 || slot >= KID_MAX_SLOTS) return fail(KID_EINVAL, "kid_wait: bad sample or slot");
    DeviceGuard guard(s->db->device);
    if (s->slot[slot].stream) KID_CUDA(cudaStreamSynchronize(s->slot[slot].stream));
    return KID_OK;
}

// ---------------------------------------------------------------------------------- sample end
int kid_sample_counts(kid_sample *s, int32_t *gcount, int32_t *ucount, void *stream_)
{
    if (!s) return fail(KID_EINVAL, "kid_sample_counts: s is NULL");
    DeviceGuard guard(s->db->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t nb = sizeof(int) * (size_t)s->db->n_taxa;
    if (ucount) {
        KID_CUDA(cudaMemsetAsync(s->ucount, 0, nb, stream));
        KID_CUDA(kid_launch_ucount(s->db->table_ptr(), s->db->layout, s->seen, 0, s->n_words, s->ucount,
                                   s->db->n_taxa, stream));
        KID_CUDA(cudaMemcpyAsync(ucount, s->ucount, nb, cudaMemcpyDeviceToHost, stream));
    }
    if (gcount) KID_CUDA(cudaMemcpyAsync(gcount, s->gcount, nb, cudaMemcpyDeviceToHost, stream));
    KID_CUDA(cudaStreamSynchronize(stream));
    s->d2h += (gcount ? nb : 0) + (ucount ? nb : 0);
    return KID_OK;
}

int kid_samples_counts(kid_sample *const *samples, int n, int32_t *gcount, int32_t *ucount, void *stream_)
{
    if (!samples || n < 1 || n > KID_MAX_OR_SOURCES) return fail(KID_EINVAL, "kid_samples_counts: bad argument");
    for (int i = 0; i < n; i++)
        if (!samples[i] || samples[i]->db != samples[0]->db)
            return fail(KID_EINVAL, "kid_samples_counts: samples must share one kid_db");
    kid_sample *s0 = samples[0];
    DeviceGuard guard(s0->db->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t nt = (size_t)s0->db->n_taxa, nb = sizeof(int) * nt;
    if (ucount) {
        KidPtrList l;
        for (int i = 0; i < KID_MAX_OR_SOURCES; i++) l.p[i] = i < n ? samples[i]->seen : nullptr;
        KID_CUDA(cudaMemsetAsync(s0->ucount, 0, nb, stream));
        KID_CUDA(kid_launch_ucount_or(s0->db->table_ptr(), s0->db->layout, l, n, 0, s0->n_words, s0->ucount,
                                      s0->db->n_taxa, stream));
        KID_CUDA(cudaMemcpyAsync(ucount, s0->ucount, nb, cudaMemcpyDeviceToHost, stream));
    }
    if (gcount) {
        std::vector<int32_t> tmp(nt);
        memset(gcount, 0, nb);
        for (int i = 0; i < n; i++) {
            KID_CUDA(cudaMemcpyAsync(tmp.data(), samples[i]->gcount, nb, cudaMemcpyDeviceToHost, stream));
            KID_CUDA(cudaStreamSynchronize(stream));
            for (size_t t = 0; t < nt; t++) gcount[t] += tmp[t];
        }
    }
    KID_CUDA(cudaStreamSynchronize(stream));
    return KID_OK;
}

int kid_sample_counters(kid_sample *s, uint64_t *lookups, uint64_t *hits, uint64_t *reads, void *stream_)
{
    if (!s) return fail(KID_EINVAL, "kid_sample_counters: s is NULL");
    DeviceGuard guard(s->db->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    unsigned long long c[2];
    KID_CUDA(cudaMemcpyAsync(c, s->counters, sizeof c, cudaMemcpyDeviceToHost, stream));
    if (reads) {
        std::vector<int> g((size_t)s->db->n_taxa);
        KID_CUDA(cudaMemcpyAsync(g.data(), s->gcount, sizeof(int) * g.size(), cudaMemcpyDeviceToHost, stream));
        KID_CUDA(cudaStreamSynchronize(stream));
        uint64_t t = 0;
        for (int v : g) t += (uint64_t)v;
        *reads = t;
    }
    KID_CUDA(cudaStreamSynchronize(stream));
    if (lookups) *lookups = c[0];
    if (hits) *hits = c[1];
    return KID_OK;
}

int kid_sample_gcount_device(kid_sample *s, int32_t **gcount)
{
    if (!s || !gcount) return fail(KID_EINVAL, "kid_sample_gcount_device: NULL argument");
    *gcount = s->gcount;
    return KID_OK;
}

int kid_sample_seen_device(kid_sample *s, uint32_t **seen, uint64_t *n_words)
{
    if (!s || !seen) return fail(KID_EINVAL, "kid_sample_seen_device: NULL argument");
    *seen = s->seen;
    if (n_words) *n_words = s->n_words;
    return KID_OK;
}

int kid_peer_enable(const int *devices, int n)
{
    if (!devices || n < 1) return fail(KID_EINVAL, "kid_peer_enable: bad argument");
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            if (i == j || devices[i] == devices[j]) continue;
            int can = 0;
            KID_CUDA(cudaDeviceCanAccessPeer(&can, devices[i], devices[j]));
            if (!can) return fail(KID_ECUDA, "device %d cannot map the memory of device %d", devices[i], devices[j]);
            DeviceGuard guard(devices[i]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); continue; }
            if (e != cudaSuccess) return fail(KID_ECUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", devices[i], devices[j], cudaGetErrorString(e));
        }
    return KID_OK;
}

int kid_sample_ucount_partial(kid_sample *s, kid_sample *const *shards, int n_shards, int part, int n_parts, void *stream_)
{
    if (!s || !shards || n_shards < 1 || n_shards > KID_MAX_OR_SOURCES || n_parts < 1 || part < 0 || part >= n_parts)
        return fail(KID_EINVAL, "kid_sample_ucount_partial: bad argument");
    KidPtrList l;
    for (int i = 0; i < KID_MAX_OR_SOURCES; i++) l.p[i] = nullptr;
    for (int i = 0; i < n_shards; i++) {
        const kid_sample *t = shards[i];
        if (!t) return fail(KID_EINVAL, "kid_sample_ucount_partial: shard %d is NULL", i);
        const kid_db *a = s->db, *b = t->db;
        // slot i must mean the same key on every replica
        if (a->layout != b->layout || a->log2_sectors != b->log2_sectors || a->sub_bits != b->sub_bits || a->mm != b->mm ||
            a->n_distinct != b->n_distinct || a->n_displaced != b->n_displaced || a->n_taxa != b->n_taxa ||
            s->n_words != t->n_words)
            return fail(KID_EINVAL, "kid_sample_ucount_partial: shard %d sits on a table replica of another shape "
                                    "(2^%d vs 2^%d sectors, %llu vs %llu keys)", i, b->log2_sectors, a->log2_sectors,
                        (unsigned long long)b->n_distinct, (unsigned long long)a->n_distinct);
        l.p[i] = t->seen;
    }
    // equal word ranges, multiples of 4 words; the last part takes the remainder
    const uint64_t per = (s->n_words / (uint64_t)n_parts) & ~(uint64_t)3;
    const uint64_t w0 = per * (uint64_t)part;
    const uint64_t nw = part == n_parts - 1 ? s->n_words - w0 : per;
    DeviceGuard guard(s->db->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    KID_CUDA(cudaMemsetAsync(s->ucount, 0, sizeof(int) * (size_t)s->db->n_taxa, stream));
    if (nw)
        KID_CUDA(kid_launch_ucount_or(s->db->table_ptr(), s->db->layout, l, n_shards, w0, nw, s->ucount, s->db->n_taxa, stream));
    return KID_OK;
}

int kid_sample_ucount_device(kid_sample *s, int32_t **ucount)
{
    if (!s || !ucount) return fail(KID_EINVAL, "kid_sample_ucount_device: NULL argument");
    *ucount = s->ucount;
    return KID_OK;
}

int kid_sample_read_counts(kid_sample *s, int32_t *gcount, int32_t *ucount, void *stream_)
{
    if (!s) return fail(KID_EINVAL, "kid_sample_read_counts: s is NULL");
    DeviceGuard guard(s->db->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t nb = sizeof(int) * (size_t)s->db->n_taxa;
    if (gcount) KID_CUDA(cudaMemcpyAsync(gcount, s->gcount, nb, cudaMemcpyDeviceToHost, stream));
    if (ucount) KID_CUDA(cudaMemcpyAsync(ucount, s->ucount, nb, cudaMemcpyDeviceToHost, stream));
    KID_CUDA(cudaStreamSynchronize(stream));
    return KID_OK;
}

int kid_device_sync(int device)
{
    DeviceGuard guard(device);
    if (!guard.ok) return fail(KID_ECUDA, "cudaSetDevice(%d) failed", device);
    KID_CUDA(cudaDeviceSynchronize());
    return KID_OK;
}

int kid_seen_or_device(const kid_db *db, uint32_t *dst, const uint32_t *const *src, int n_src,
                       uint64_t word0, uint64_t n_words, void *stream)
{
    if (!db || !dst || !src) return fail(KID_EINVAL, "kid_seen_or_device: NULL argument");
    if (n_src < 1 || n_src > KID_MAX_OR_SOURCES)
        return fail(KID_EINVAL, "kid_seen_or_device: n_src %d outside [1,%d]", n_src, KID_MAX_OR_SOURCES);
    if ((word0 | n_words) & 3 || (reinterpret_cast<uintptr_t>(dst) & 15))
        return fail(KID_EINVAL, "kid_seen_or_device: ranges must be multiples of 4 words, dst 16-byte aligned");
    KidPtrList l;
    for (int i = 0; i < KID_MAX_OR_SOURCES; i++) l.p[i] = i < n_src ? src[i] : nullptr;
    DeviceGuard guard(db->device);
    KID_CUDA(kid_launch_seen_or(dst, l, n_src, word0, n_words, (cudaStream_t)stream));
    return KID_OK;
}

int kid_sample_use_seen_buffer(kid_sample *s, uint32_t *buf, uint64_t n_words)
{
    if (!s) return fail(KID_EINVAL, "kid_sample_use_seen_buffer: s is NULL");
    if (!buf) { s->seen = s->own_seen; return KID_OK; }
    if (n_words < s->n_words || (reinterpret_cast<uintptr_t>(buf) & 15))
        return fail(KID_EINVAL, "kid_sample_use_seen_buffer: need >= %llu words, 16-byte aligned",
                    (unsigned long long)s->n_words);
    s->seen = buf;
    return KID_OK;
}

int kid_ucount_or_range_device(const kid_db *db, const uint32_t *const *seen_srcs, int n_src, uint64_t word0,
                               uint64_t n_words, int32_t *ucount_partial, void *stream)
{
    if (!db || !seen_srcs || !ucount_partial) return fail(KID_EINVAL, "kid_ucount_or_range_device: NULL argument");
    if (n_src < 1 || n_src > KID_MAX_OR_SOURCES)
        return fail(KID_EINVAL, "kid_ucount_or_range_device: n_src %d outside [1,%d]", n_src, KID_MAX_OR_SOURCES);
    if ((word0 | n_words) & 3) return fail(KID_EINVAL, "kid_ucount_or_range_device: range must be multiples of 4 words");
    KidPtrList l;
    for (int i = 0; i < KID_MAX_OR_SOURCES; i++) l.p[i] = i < n_src ? seen_srcs[i] : nullptr;
    DeviceGuard guard(db->device);
    KID_CUDA(kid_launch_ucount_or(db->table_ptr(), db->layout, l, n_src, word0, n_words, ucount_partial, db->n_taxa,
                                  (cudaStream_t)stream));
    return KID_OK;
}

int kid_ucount_range_device(const kid_db *db, const uint32_t *seen, uint64_t word0, uint64_t n_words,
                            int32_t *ucount_partial, void *stream)
{
    if (!db || !seen || !ucount_partial) return fail(KID_EINVAL, "kid_ucount_range_device: NULL argument");
    if ((word0 | n_words) & 3) return fail(KID_EINVAL, "kid_ucount_range_device: range must be multiples of 4 words");
    DeviceGuard guard(db->device);
    KID_CUDA(kid_launch_ucount(db->table_ptr(), db->layout, seen, word0, n_words, ucount_partial, db->n_taxa,
                               (cudaStream_t)stream));
    return KID_OK;
}

} // extern "C"
