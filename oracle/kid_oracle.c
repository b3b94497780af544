/*
 * kid_oracle.c - CPU restatement of the nk10 read-classification path (see kid_oracle.h).
 * TEST INFRASTRUCTURE ONLY - never linked into the product library.
 *
 * Citations are into /root/reference/newkmer_10nx.cpp unless another file is named.
 * The data structures are deliberately different from the reference's (a linear-probing
 * first-wins map with a per-entry "seen" flag instead of a 2^30-cell triangular-probing table
 * plus a std::set) - only the observable results are restated.
 */
#define _POSIX_C_SOURCE 200809L
#include "kid_oracle.h"

#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define KOR_ROOT 1
#define KOR_SAVENUM 12         /* :48 */
#define KOR_BUFLEN 0x4000      /* :85 - a line of >= 16384 bytes makes the reference exit(255) */
#define KOR_MASK60 ((1ULL << 60) - 1) /* :76 */

typedef struct {
    uint64_t key; /* key + 1 so that 0 means empty (keys are < 2^60) */
    uint32_t value;
    uint32_t slot_id; /* dense id of this entry, index into kor_sample.seen */
} kor_cell;

struct kor_db {
    int n_taxa;
    unsigned flags;
    int *parent;
    kor_cell *cells;
    uint64_t cap; /* power of two */
    uint64_t n_keys;
};

struct kor_sample {
    const kor_db *db;
    int32_t *gcount, *ucount;
    uint8_t *seen; /* by slot_id; equivalent to the reference's std::set<ktype> kmer_seen (:64) */
    uint64_t seen_cap;
    uint64_t lookups, hits;
    long long tct;
};

/* ---------------------------------------------------------------------------------------- */
static uint64_t mix64(uint64_t k) /* any good mixer will do; the map is semantic, not layout */
{
    k ^= k >> 31;
    k *= 0x9E3779B97F4A7C15ULL;
    k ^= k >> 29;
    k *= 0xBF58476D1CE4E5B9ULL;
    k ^= k >> 32;
    return k;
}

kor_db *kor_db_new(int n_taxa, unsigned flags)
{
    kor_db *db = (kor_db *)calloc(1, sizeof *db);
    if (!db || n_taxa < 2) { free(db); return NULL; }
    db->n_taxa = n_taxa;
    db->flags = flags;
    db->parent = (int *)malloc(sizeof(int) * (size_t)n_taxa);
    for (int i = 0; i < n_taxa; i++) db->parent[i] = KOR_ROOT; /* Tree1():101-106 */
    db->cap = 1u << 16;
    db->cells = (kor_cell *)calloc(db->cap, sizeof(kor_cell));
    return db;
}

void kor_db_free(kor_db *db)
{
    if (!db) return;
    free(db->parent);
    free(db->cells);
    free(db);
}

int kor_db_n_taxa(const kor_db *db) { return db->n_taxa; }
uint64_t kor_db_n_keys(const kor_db *db) { return db->n_keys; }

/* Tree1::add_edge :112-116 (children[] is never read on this path) */
int kor_db_add_edge(kor_db *db, int parent, int child)
{
    if (parent < 0 || parent >= db->n_taxa || child < 0 || child >= db->n_taxa) return -1;
    db->parent[child] = parent;
    return 0;
}

/* Tree1::get_parent :146-152 */
int kor_db_parent(const kor_db *db, int x)
{
    if (x != KOR_ROOT && x > 0) return db->parent[x];
    return KOR_ROOT;
}

/* membership in the set built at :124-132 = {root} U {x, gp(x), ... up to but excluding root} */
static int in_ancestors(const kor_db *db, int x, int z)
{
    if (z == KOR_ROOT) return 1;
    for (int w = x; w != KOR_ROOT; w = kor_db_parent(db, w))
        if (w == z) return 1;
    return 0;
}

/* Tree1::msca :118-144 */
int kor_msca(const kor_db *db, int x, int y)
{
    int z = y;
    if (in_ancestors(db, x, y)) return x; /* :134-135 */
    while (!in_ancestors(db, x, z)) {      /* :136 */
        z = kor_db_parent(db, z);          /* :138 */
        if (z == x) return y;              /* :139-140 */
    }
    return z; /* :143 */
}

/* ---- istream-style extraction (what `ss >> x` does in the C locale, C++11 rules) ---------- */
static int is_ws(unsigned char c)
{
    return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r';
}

typedef struct {
    const char *p, *end;
    int fail;
} kor_cur;

static void skip_ws(kor_cur *c)
{
    while (c->p < c->end && is_ws((unsigned char)*c->p)) c->p++;
}

/* magnitude + sign; returns 0 if no digits.  *ovf set if magnitude exceeds 2^64-1 */
static int scan_int(kor_cur *c, int *neg, unsigned long long *mag, int *ovf)
{
    skip_ws(c);
    *neg = 0; *mag = 0; *ovf = 0;
    const char *q = c->p;
    if (q < c->end && (*q == '+' || *q == '-')) { *neg = (*q == '-'); q++; }
    const char *d0 = q;
    while (q < c->end && *q >= '0' && *q <= '9') {
        unsigned d = (unsigned)(*q - '0');
        if (*mag > (ULLONG_MAX - d) / 10) *ovf = 1;
        else *mag = *mag * 10ULL + d;
        q++;
    }
    if (q == d0) return 0; /* nothing consumed counts as failure; libstdc++ leaves the sign eaten */
    c->p = q;
    return 1;
}

static void get_int(kor_cur *c, int *out) /* istream >> int */
{
    int neg, ovf; unsigned long long mag;
    if (c->fail) return;
    if (!scan_int(c, &neg, &mag, &ovf)) { *out = 0; c->fail = 1; return; }
    if (ovf || (!neg && mag > (unsigned long long)INT_MAX) || (neg && mag > (unsigned long long)INT_MAX + 1ULL)) {
        *out = neg ? INT_MIN : INT_MAX; c->fail = 1; return;
    }
    *out = neg ? (int)(-(long long)mag) : (int)mag;
}

static void get_uint(kor_cur *c, unsigned *out) /* istream >> unsigned int */
{
    int neg, ovf; unsigned long long mag;
    if (c->fail) return;
    if (!scan_int(c, &neg, &mag, &ovf)) { *out = 0; c->fail = 1; return; }
    if (ovf || mag > (unsigned long long)UINT_MAX) { *out = UINT_MAX; c->fail = 1; return; }
    *out = neg ? (unsigned)(0u - (unsigned)mag) : (unsigned)mag;
}

/* main():976-981  `linestream >> i >> j; add_edge(i,j)` */
int kor_db_add_edge_line(kor_db *db, const char *line, size_t len, int state[2])
{
    kor_cur c = { line, line + len, 0 };
    get_int(&c, &state[0]);
    get_int(&c, &state[1]);
    return kor_db_add_edge(db, state[0], state[1]);
}

int kor_db_load_tree(kor_db *db, const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) return -1; /* :974 - silently skipped by the reference */
    char *line = NULL; size_t cap = 0; ssize_t n; int st[2] = { 0, 0 }; int rc = 0;
    while ((n = getline(&line, &cap, f)) >= 0) {
        if (n > 0 && line[n - 1] == '\n') n--;
        if (kor_db_add_edge_line(db, line, (size_t)n, st) != 0) rc = -2;
    }
    free(line);
    fclose(f);
    return rc;
}

/* ---- first-wins map (semantics of add_kmer :235-263 + getHash :204-233) -------- */
static void db_grow(kor_db *db)
{
    uint64_t ncap = db->cap * 2;
    kor_cell *nc = (kor_cell *)calloc(ncap, sizeof(kor_cell));
    for (uint64_t i = 0; i < db->cap; i++) {
        if (!db->cells[i].key) continue;
        uint64_t j = mix64(db->cells[i].key - 1) & (ncap - 1);
        while (nc[j].key) j = (j + 1) & (ncap - 1);
        nc[j] = db->cells[i];
    }
    free(db->cells);
    db->cells = nc;
    db->cap = ncap;
}

void kor_db_add_key(kor_db *db, uint64_t key, uint32_t target)
{
    if (target == 0) return; /* value==0 is "empty" (:219,:248): such a line is never visible */
    if ((db->n_keys + 1) * 2 > db->cap) db_grow(db);
    uint64_t j = mix64(key) & (db->cap - 1);
    while (db->cells[j].key) {
        if (db->cells[j].key == key + 1) return; /* an earlier line owns this key: it shadows us */
        j = (j + 1) & (db->cap - 1);
    }
    db->cells[j].key = key + 1;
    db->cells[j].value = target;
    db->cells[j].slot_id = (uint32_t)db->n_keys++;
}

void kor_db_reserve(kor_db *db, uint64_t n_keys)
{
    while (db->cap < 2 * n_keys + 2) db_grow(db);
}

void kor_db_add_keys(kor_db *db, const uint64_t *keys, const uint32_t *taxa, size_t n)
{
    kor_db_reserve(db, db->n_keys + n);
    for (size_t i = 0; i < n; i++) kor_db_add_key(db, keys[i], taxa[i]);
}

static const kor_cell *db_find(const kor_db *db, uint64_t key)
{
    uint64_t j = mix64(key) & (db->cap - 1);
    while (db->cells[j].key) {
        if (db->cells[j].key == key + 1) return &db->cells[j];
        j = (j + 1) & (db->cap - 1);
    }
    return NULL;
}

uint32_t kor_db_lookup(const kor_db *db, uint64_t key)
{
    const kor_cell *c = db_find(db, key);
    return c ? c->value : 0;
}

/* process_kmer :619-661 - upper-case ACGT only, forward strand only, every full window */
int kor_db_add_probe_seq(kor_db *db, const char *seq, size_t len, uint32_t target)
{
    int cpos = 0, n = 0;
    uint64_t kf = 0;
    for (size_t i = 0; i < len; i++) {
        int c;
        switch (seq[i]) {
        case 'A': c = 0; break;
        case 'C': c = 1; break;
        case 'G': c = 2; break;
        case 'T': c = 3; break;
        default: c = -1; break;
        }
        if (c < 0) { cpos = 0; kf = 0; continue; } /* :649-652 */
        kf = ((kf << 2) & KOR_MASK60) | (uint64_t)c;
        if (++cpos == KOR_KSIZE) { /* :654-658 */
            kor_db_add_key(db, kf, target);
            n++;
            cpos--;
        }
    }
    return n;
}

/* process_kmergz :690-702 for one line (already split on '\n') */
int kor_db_add_probe_line(kor_db *db, const char *line, size_t len)
{
    if (len > 0 && line[len - 1] == '\r') len--; /* :691-692 */
    if (len == 0) return 0;                       /* :693 */
    char stackbuf[512];
    char *tmp = len < sizeof stackbuf ? stackbuf : (char *)malloc(len + 1);
    for (size_t i = 0; i < len; i++) tmp[i] = line[i] == ',' ? ' ' : line[i]; /* :695 */
    kor_cur c = { tmp, tmp + len, 0 };
    /* ss >> sequence */
    skip_ws(&c);
    const char *s0 = c.p;
    while (c.p < c.end && !is_ws((unsigned char)*c.p)) c.p++;
    size_t slen = (size_t)(c.p - s0);
    int ok = 0;
    if (slen > 0) {
        unsigned target = 0; int org = 0, position = 0, count = 0;
        get_uint(&c, &target);
        get_int(&c, &org);
        get_int(&c, &position);
        if (!c.fail) { /* >> strand (char) */
            skip_ws(&c);
            if (c.p < c.end) c.p++; else c.fail = 1;
        }
        get_int(&c, &count);
        if (!c.fail) { /* :697-701 */
            kor_db_add_probe_seq(db, s0, slen, target);
            ok = 1;
        }
    }
    if (tmp != stackbuf) free(tmp);
    return ok;
}

/* Splits a gz stream into '\n'-terminated lines exactly like the 16 KiB loop shared by
 * process_kmergz :675-707 and process_fqgz :770-810: an unterminated tail is dropped, a line of
 * >= BUFLEN bytes is the "Buffer to small" exit(255). */
typedef int (*kor_line_fn)(void *ctx, const char *line, size_t len);

static int for_each_gz_line(const char *path, kor_line_fn fn, void *ctx)
{
    gzFile in = gzopen(path, "rb");
    if (!in) return -255; /* gzread(NULL) < 0 -> error() -> exit(255) */
    gzbuffer(in, 1 << 20);
    size_t cap = 1 << 22, have = 0;
    char *buf = (char *)malloc(cap);
    int rc = 0;
    for (;;) {
        int got = gzread(in, buf + have, (unsigned)(cap - have));
        if (got < 0) { rc = -255; break; }
        if (got == 0) break;
        size_t end = have + (size_t)got, cur = 0;
        while (cur < end) {
            char *eol = (char *)memchr(buf + cur, '\n', end - cur);
            if (!eol) break;
            size_t len = (size_t)(eol - (buf + cur));
            if (len >= KOR_BUFLEN) { rc = -255; goto done; }
            rc = fn(ctx, buf + cur, len);
            if (rc) goto done;
            cur += len + 1;
        }
        have = end - cur;
        if (have >= KOR_BUFLEN) { rc = -255; break; }
        memmove(buf, buf + cur, have);
    }
done:
    free(buf);
    if (gzclose(in) != Z_OK && rc == 0) rc = -255; /* :710,:815 "failed gzclose" */
    return rc;
}

typedef struct { kor_db *db; long long lines; } probe_ctx;

static int probe_line_cb(void *vctx, const char *line, size_t len)
{
    probe_ctx *p = (probe_ctx *)vctx;
    p->lines += kor_db_add_probe_line(p->db, line, len);
    return 0;
}

long long kor_db_load_probes_gz(kor_db *db, const char *path)
{
    probe_ctx p = { db, 0 };
    int rc = for_each_gz_line(path, probe_line_cb, &p);
    return rc ? rc : p.lines;
}

/* ---- per-read pieces --------------------------------------------------------------------- */
/* process_qual :724-753.  `char` is signed on the reference's platform, so bytes >= 0x80 count as
 * very low quality; window sums are plain int arithmetic. */
void kor_trim(const char *qual_, int seqlen, int *start_out, int *stop_out)
{
    const signed char *qual = (const signed char *)qual_;
    const int cutoff_char = 32 + 17, window_size = 4, window_cut = 17 * 4;
    int stop = seqlen - 1, start = 0, window_val, i;
    while (qual[start] < cutoff_char && start < stop) start++;
    while (qual[stop] < cutoff_char && stop > start) stop--;
    if (start < stop - window_size) {
        window_val = 0;
        for (i = 0; i < window_size; i++) window_val += qual[start + i] - 32;
        while (window_val < window_cut && start < stop - window_size) {
            window_val += qual[start + window_size] - qual[start];
            start++;
        }
    }
    if (start < stop - window_size) {
        window_val = 0;
        for (i = 0; i < window_size; i++) window_val += qual[stop - i] - 32;
        while (window_val < window_cut && start < stop - window_size) {
            window_val += qual[stop - window_size] - qual[stop];
            stop--;
        }
    }
    *start_out = start;
    *stop_out = stop;
}

static int base_code(char b, unsigned flags) /* :478-525; vf6 adds U/u kmer_read_vf6.cpp:496-525 */
{
    switch (b) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    case 'U': case 'u': return (flags & KOR_FLAG_ACCEPT_U) ? 3 : -1;
    default: return -1;
    }
}

int kor_canonical_key(const char *seq, unsigned flags, uint64_t *key)
{
    uint64_t kf = 0, kr = 0;
    for (int i = 0; i < KOR_KSIZE; i++) {
        int c = base_code(seq[i], flags);
        if (c < 0) return 0;
        kf = ((kf << 2) & KOR_MASK60) | (uint64_t)c;
        kr = (kr >> 2) | ((uint64_t)(3 - c) << 58);
    }
    *key = kf < kr ? kf : kr;
    return 1;
}

/* ---- sample accumulators ------------------------------------------------------------------ */
kor_sample *kor_sample_new(const kor_db *db)
{
    kor_sample *s = (kor_sample *)calloc(1, sizeof *s);
    s->db = db;
    s->gcount = (int32_t *)calloc((size_t)db->n_taxa, sizeof(int32_t));
    s->ucount = (int32_t *)calloc((size_t)db->n_taxa, sizeof(int32_t));
    s->seen_cap = db->n_keys ? db->n_keys : 1;
    s->seen = (uint8_t *)calloc(s->seen_cap, 1);
    return s;
}

void kor_sample_free(kor_sample *s)
{
    if (!s) return;
    free(s->gcount); free(s->ucount); free(s->seen); free(s);
}

void kor_sample_reset(kor_sample *s) /* :1017-1023 */
{
    memset(s->gcount, 0, sizeof(int32_t) * (size_t)s->db->n_taxa);
    memset(s->ucount, 0, sizeof(int32_t) * (size_t)s->db->n_taxa);
    if (s->seen_cap < s->db->n_keys) { /* db grew after the sample was made */
        free(s->seen);
        s->seen_cap = s->db->n_keys;
        s->seen = (uint8_t *)calloc(s->seen_cap, 1);
    } else {
        memset(s->seen, 0, s->seen_cap);
    }
    s->lookups = s->hits = 0;
    s->tct = 0;
}

const int32_t *kor_sample_gcount(const kor_sample *s) { return s->gcount; }
const int32_t *kor_sample_ucount(const kor_sample *s) { return s->ucount; }
uint64_t kor_sample_lookups(const kor_sample *s) { return s->lookups; }
uint64_t kor_sample_hits(const kor_sample *s) { return s->hits; }
long long kor_sample_tct(const kor_sample *s) { return s->tct; }

/* process_read :452-617 (the Smith-Waterman branch :530-587 is dead: minalign = 0, :27) */
int kor_process_read(kor_sample *s, const char *seq, int start, int stop, const char *acc,
                     size_t acclen, FILE *reads_out)
{
    const kor_db *db = s->db;
    uint64_t kf = 0, kr = 0;
    int cpos = 0;
    uint32_t final_targ = 0;
    for (int i = start; i <= stop; i++) {
        int c = base_code(seq[i], db->flags);
        if (c < 0) { cpos = 0; kf = 0; kr = 0; continue; } /* :520-524 */
        kf = ((kf << 2) & KOR_MASK60) | (uint64_t)c;          /* :481 etc. */
        kr = (kr >> 2) | ((uint64_t)(3 - c) << 58);           /* :482 etc. */
        cpos++;
        if (cpos == KOR_KSIZE) {                              /* :526 */
            uint64_t key = kf < kr ? kf : kr;                 /* :528 */
            const kor_cell *cell = db_find(db, key);          /* :529 */
            uint32_t target = cell ? cell->value : 0;
            s->lookups++;
            if (target > 0) s->hits++;
            if (final_targ > 0 && target > 0)                 /* :588-595 */
                final_targ = (uint32_t)kor_msca(db, (int)target, (int)final_targ);
            else if (target > 0)
                final_targ = target;
            if (target > 1 && !s->seen[cell->slot_id]) {      /* :596-603 */
                s->seen[cell->slot_id] = 1;
                s->ucount[target]++;
            }
            cpos--;                                           /* :604 */
        }
    }
    if (reads_out && final_targ > 1 && s->gcount[final_targ] < KOR_SAVENUM) { /* :608-612 */
        fprintf(reads_out, ">%u:", final_targ);
        fwrite(acc, 1, acclen, reads_out);
        fputc('\n', reads_out);
        fwrite(seq + start, 1, (size_t)(stop - start + 1), reads_out);
        fputc('\n', reads_out);
    }
    s->gcount[final_targ]++; /* :613 */
    s->tct++;                /* :614 */
    return (int)final_targ;
}

int kor_process_qual(kor_sample *s, const char *seq, int seqlen, const char *qual, int quallen,
                     const char *acc, size_t acclen, FILE *reads_out, int *start_out,
                     int *stop_out)
{
    int start, stop;
    if (quallen < seqlen) return -2; /* qual.at(stop) throws std::out_of_range (:729) -> abort */
    kor_trim(qual, seqlen, &start, &stop);
    if (start_out) *start_out = start;
    if (stop_out) *stop_out = stop;
    if (stop - start >= KOR_KSIZE) /* :755 */
        return kor_process_read(s, seq, start, stop, acc, acclen, reads_out);
    return -1;
}

void kor_classify_batch(kor_sample *s, const char *seq, const char *qual, const uint64_t *off,
                        size_t n, int32_t *out_final, int32_t *out_start, int32_t *out_stop)
{
    for (size_t r = 0; r < n; r++) {
        int len = (int)(off[r + 1] - off[r]);
        int start = 0, stop = len - 1, fin = -1;
        if (len > 0) {
            if (qual) {
                fin = kor_process_qual(s, seq + off[r], len, qual + off[r], len, "", 0, NULL,
                                       &start, &stop);
            } else if (len > KOR_KSIZE) { /* process_fagz :849-852 */
                fin = kor_process_read(s, seq + off[r], 0, len - 1, "", 0, NULL);
            }
        }
        if (out_final) out_final[r] = fin;
        if (out_start) out_start[r] = start;
        if (out_stop) out_stop[r] = stop;
    }
}

/* ---- process_fqgz :762-816 ------------------------------------------------------------------ */
typedef struct {
    kor_sample *s;
    FILE *reads_out;
    int mod4;
    char *acc, *seq;
    size_t acclen, seqlen;
} fq_ctx;

static int fq_line_cb(void *vctx, const char *line, size_t len)
{
    fq_ctx *f = (fq_ctx *)vctx;
    if (len > 0 && line[len - 1] == '\r') len--; /* :786-787 */
    if (len == 0) return 0;                       /* :788 - empty lines do not advance mod4 */
    if (f->mod4 == 1) {
        memcpy(f->seq, line, len); f->seqlen = len;
    } else if (f->mod4 == 0) {
        memcpy(f->acc, line, len); f->acclen = len;
    } else if (f->mod4 == 3) {
        int rc = kor_process_qual(f->s, f->seq, (int)f->seqlen, line, (int)len, f->acc, f->acclen,
                                  f->reads_out, NULL, NULL);
        if (rc == -2) return -134;
    }
    f->mod4 = (f->mod4 + 1) % 4;
    return 0;
}

int kor_run_fastq_gz(kor_sample *s, const char *path, FILE *reads_out)
{
    fq_ctx f = { s, reads_out, 0, NULL, NULL, 0, 0 };
    f.acc = (char *)malloc(KOR_BUFLEN);
    f.seq = (char *)malloc(KOR_BUFLEN);
    int rc = for_each_gz_line(path, fq_line_cb, &f);
    free(f.acc);
    free(f.seq);
    return rc;
}

int kor_write_result(const kor_sample *s, const char *path) /* :1040-1043 */
{
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    for (int i = 0; i < s->db->n_taxa; i++) fprintf(f, "%d,%d,%d\n", i, s->gcount[i], s->ucount[i]);
    fclose(f);
    return 0;
}
