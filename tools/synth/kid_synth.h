/*
 * kid_synth.h - counter-based synthetic workload shared by the CPU and the GPU generators.
 *
 * BENCH / TEST TOOLING, not part of the product library.  Implements the workload SURVEY.md
 * section 8(d) specifies for BASELINE.json's configs: a probe database whose per-taxon probe counts
 * come from the shipped refkey (random canonical 30-mers tagged with b10 taxa, taxon-sorted file
 * order) and reads stitched from whole probes of one lineage (leaf + ancestors) with random
 * spacers, substitutions, N's and low-quality tails.  Everything is a pure function of
 * (seed, index), so any slice can be produced on any device and the CPU and GPU generators agree
 * bit for bit (tests/test_synth.py).
 */
#ifndef KID_SYNTH_H
#define KID_SYNTH_H
#include <stddef.h>
#include <stdint.h>

#ifdef __CUDACC__
#define KS_HD __host__ __device__ __forceinline__
#else
#define KS_HD static inline
#endif

#define KS_MASK60 ((1ULL << 60) - 1ULL)
#define KS_MAX_PATH 12

typedef struct {
    uint64_t seed_db;
    uint64_t seed_reads;
    uint64_t n_probes;       /* prefix[n_taxa] */
    uint32_t n_taxa;
    uint32_t read_len;       /* bases per read */
    uint32_t stride;         /* bytes between read starts in the output buffers (>= read_len) */
    uint32_t on_target_pct;  /* 70 */
    uint32_t sub_per_10k;    /* 50  = 0.5 % substitutions */
    uint32_t n_per_10k;      /* 10  = 0.1 % N */
    uint32_t bad_tail_pct;   /* 20 % of reads get a 1..25-base tail decaying to '#' */
    uint32_t reserved;
} ks_config;

KS_HD uint64_t ks_mix(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

KS_HD uint64_t ks_rand(uint64_t seed, uint64_t a, uint64_t b)
{
    return ks_mix(ks_mix(seed ^ ks_mix(a)) + b);
}

KS_HD uint64_t ks_revcomp60(uint64_t k)
{
    uint64_t r = 0;
    for (int i = 0; i < 30; i++) {
        r = (r << 2) | (3ULL - (k & 3ULL));
        k >>= 2;
    }
    return r;
}

/* canonical key of probe line i (what the builder writes: kmer_build_vf6.cpp:604,622) */
KS_HD uint64_t ks_probe_key(uint64_t seed_db, uint64_t i)
{
    const uint64_t k = ks_rand(seed_db, 1, i) & KS_MASK60;
    const uint64_t r = ks_revcomp60(k);
    return k < r ? k : r;
}

/* taxon owning probe index i: prefix[t] <= i < prefix[t+1] */
KS_HD uint32_t ks_taxon_of(const uint64_t *prefix, uint32_t n_taxa, uint64_t i)
{
    uint32_t lo = 0, hi = n_taxa;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (prefix[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

/* One read: g = global read index (pair = g >> 1, mate = g & 1).  Writes read_len bytes. */
KS_HD void ks_make_read(const ks_config *c, const uint64_t *prefix, const int32_t *parent,
                        uint64_t g, uint8_t *seq, uint8_t *qual)
{
    const char B[4] = { 'A', 'C', 'G', 'T' };
    const uint64_t pair = g >> 1;
    const uint32_t L = c->read_len;
    uint64_t ctr = 0;
    /* pair-level choices */
    const uint64_t pr = ks_rand(c->seed_reads, 0x70616972ULL, pair);
    const int on_target = c->n_probes > 0 && (pr % 100) < c->on_target_pct;
    uint32_t path[KS_MAX_PATH];
    int npath = 0;
    if (on_target) {
        uint32_t t = ks_taxon_of(prefix, c->n_taxa, (pr >> 8) % c->n_probes);
        while (t != 1 && t > 0 && npath < KS_MAX_PATH) {
            if (prefix[t + 1] > prefix[t]) path[npath++] = t;
            t = (uint32_t)parent[t];
        }
    }
    uint32_t pos = 0;
    if (npath > 0) {
        uint32_t skip = (uint32_t)(ks_rand(c->seed_reads, g + 1, ctr++) % 30);
        while (pos < L) {
            const uint64_t r = ks_rand(c->seed_reads, g + 1, ctr++);
            const uint32_t t = path[r % (uint64_t)npath];
            const uint64_t cnt = prefix[t + 1] - prefix[t];
            uint64_t k = ks_probe_key(c->seed_db, prefix[t] + (r >> 8) % cnt);
            if ((r >> 60) & 1) k = ks_revcomp60(k);
            for (int i = 29; i >= 0 && pos < L; i--) {
                if (skip) { skip--; continue; }
                seq[pos++] = (uint8_t)B[(k >> (2 * i)) & 3];
            }
            uint64_t sr = ks_rand(c->seed_reads, g + 1, ctr++);
            uint32_t sp = (uint32_t)(sr % 11);
            sr >>= 8;
            while (sp-- && pos < L) { seq[pos++] = (uint8_t)B[sr & 3]; sr >>= 2; }
        }
    } else {
        while (pos < L) {
            uint64_t r = ks_rand(c->seed_reads, g + 1, ctr++);
            for (int i = 0; i < 32 && pos < L; i++) { seq[pos++] = (uint8_t)B[r & 3]; r >>= 2; }
        }
    }
    /* sequencing errors: one 16-bit lane of a random word per base */
    for (uint32_t p0 = 0; p0 < L; p0 += 4) {
        uint64_t r = ks_rand(c->seed_reads, g + 1, 0x10000ULL + p0);
        for (uint32_t i = 0; i < 4 && p0 + i < L; i++, r >>= 16) {
            const uint32_t v = (uint32_t)(r & 0xFFFF) % 10000u;
            if (v < c->sub_per_10k) seq[p0 + i] = (uint8_t)B[(r >> 14) & 3];
            else if (v < c->sub_per_10k + c->n_per_10k) seq[p0 + i] = 'N';
        }
    }
    /* quality */
    for (uint32_t i = 0; i < L; i++) qual[i] = 'I';
    const uint64_t qr = ks_rand(c->seed_reads, g + 1, 0x20000ULL);
    if ((qr % 100) < c->bad_tail_pct) {
        uint32_t tl = 1 + (uint32_t)((qr >> 8) % 25);
        if (tl > L) tl = L;
        for (uint32_t i = 0; i < tl; i++) /* 'I' (73) down to '#' (35) */
            qual[L - tl + i] = (uint8_t)(73 - (38 * (i + 1)) / tl);
    }
}

/* ---- C-ABI of tools/libkidsynth.so ------------------------------------------------------------ */
#ifdef __cplusplus
extern "C" {
#endif
typedef struct ks_gen ks_gen;
/* prefix: n_taxa+1 host values, parent: n_taxa host values (copied) */
ks_gen *ks_create(const ks_config *cfg, const uint64_t *prefix, const int32_t *parent);
void ks_free(ks_gen *g);
/* probe entries [i0, i0+n): keys + taxa, host buffers (multi-threaded) */
int ks_db_host(ks_gen *g, uint64_t i0, uint64_t n, uint64_t *keys, uint32_t *taxa);
/* same into device buffers of `device`, on `stream` (cudaStream_t as void*) */
int ks_db_device(ks_gen *g, int device, uint64_t i0, uint64_t n, uint64_t *keys, uint32_t *taxa, void *stream);
/* reads [g0, g0+n): seq/qual buffers of n*stride bytes; bytes past read_len in each stride are
 * filled with 'N' / '!' */
int ks_reads_host(ks_gen *g, uint64_t g0, uint64_t n, uint8_t *seq, uint8_t *qual);
int ks_reads_device(ks_gen *g, int device, uint64_t g0, uint64_t n, uint8_t *seq, uint8_t *qual, void *stream);
#ifdef __cplusplus
}
#endif
#endif
