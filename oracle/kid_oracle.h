/*
 * kid_oracle.h - CPU restatement of the nk10 read-classification path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under kmer_id_b200/ may include, link or call this;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker or the timed CPU baseline.
 *
 * Every function restates one piece of /root/reference/newkmer_10nx.cpp (cited per function in
 * kid_oracle.c).  Parity pin: tests/test_oracle_vs_ref.py runs this against the compiled,
 * unmodified reference (oracle/_ref/nk10, oracle/_ref/nk10_small) on the same files and requires
 * byte-identical _result.txt and _reads.txt.  The reference ships no golden vectors of its own
 * (SURVEY.md section 8c), so the pin is differential against the reference binary.
 */
#ifndef KID_ORACLE_H
#define KID_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KOR_KSIZE 30

/* variant switches (newkmer_10nx.cpp = 0; kmer_read_vf6.cpp accepts U/u as T) */
#define KOR_FLAG_ACCEPT_U 1u

typedef struct kor_db kor_db;         /* probe table + taxonomy tree (read-only after load) */
typedef struct kor_sample kor_sample; /* gcount / ucount / seen flags for one sample */

/* ---- database ------------------------------------------------------------------------- */
kor_db *kor_db_new(int n_taxa, unsigned flags);
void kor_db_free(kor_db *db);
int kor_db_n_taxa(const kor_db *db);
uint64_t kor_db_n_keys(const kor_db *db); /* distinct visible keys */

/* Tree1::add_edge: parent[child] = parent.  Returns -1 if either id is outside [0,n_taxa). */
int kor_db_add_edge(kor_db *db, int parent, int child);
/* main():973-983 - one line of btree_10.txt with istream >> int >> int semantics (carried
 * values on a failed extraction).  state[2] holds the carried (i,j), start it at {0,0}. */
int kor_db_add_edge_line(kor_db *db, const char *line, size_t len, int state[2]);
int kor_db_load_tree(kor_db *db, const char *path); /* 0 ok / -1 file missing (silently ok in ref) */
int kor_db_parent(const kor_db *db, int x);         /* Tree1::get_parent */
int kor_msca(const kor_db *db, int x, int y);       /* Tree1::msca */

/* Hashtable::add_kmer restated as a first-wins map insert; target 0 is invisible. */
void kor_db_add_key(kor_db *db, uint64_t key, uint32_t target);
void kor_db_reserve(kor_db *db, uint64_t n_keys);
void kor_db_add_keys(kor_db *db, const uint64_t *keys, const uint32_t *taxa, size_t n); /* file order */
/* process_kmer: every forward 30-window of seq made of upper-case ACGT. Returns windows seen. */
int kor_db_add_probe_seq(kor_db *db, const char *seq, size_t len, uint32_t target);
/* one text line of probes10.txt.gz (process_kmergz:690-702); returns 1 if the line parsed. */
int kor_db_add_probe_line(kor_db *db, const char *line, size_t len);
/* whole probes10.txt.gz; returns parsed line count ("kmers loaded"), <0 on gz error. */
long long kor_db_load_probes_gz(kor_db *db, const char *path);
uint32_t kor_db_lookup(const kor_db *db, uint64_t key); /* Hashtable::getHash */

/* ---- per-read pieces ------------------------------------------------------------------ */
/* process_qual:724-753.  qual is PHRED+33 bytes compared as SIGNED char like the reference. */
void kor_trim(const char *qual, int seqlen, int *start, int *stop);
/* canonical key of the 30 bases seq[pos..pos+29]; returns 0 if any base is not ACGTacgt(Uu). */
int kor_canonical_key(const char *seq, unsigned flags, uint64_t *key);

/* ---- sample accumulators --------------------------------------------------------------- */
kor_sample *kor_sample_new(const kor_db *db);
void kor_sample_free(kor_sample *s);
void kor_sample_reset(kor_sample *s); /* main():1017-1023 */
const int32_t *kor_sample_gcount(const kor_sample *s);
const int32_t *kor_sample_ucount(const kor_sample *s);
uint64_t kor_sample_lookups(const kor_sample *s); /* number of getHash calls so far */
uint64_t kor_sample_hits(const kor_sample *s);
long long kor_sample_tct(const kor_sample *s);    /* reads processed ("reads loaded") */

/* process_read:452-617 on seq[start..stop]; returns final_targ and bumps gcount/ucount/tct.
 * If reads_out != NULL writes the _reads.txt record when the reference would (:608-612). */
int kor_process_read(kor_sample *s, const char *seq, int start, int stop, const char *acc,
                     size_t acclen, FILE *reads_out);
/* process_qual:714-760: trim then classify.  Returns final_targ, or -1 if the read is dropped
 * (trimmed span < 31), or -2 if qual is shorter than seq (the reference throws out_of_range). */
int kor_process_qual(kor_sample *s, const char *seq, int seqlen, const char *qual, int quallen,
                     const char *acc, size_t acclen, FILE *reads_out, int *start_out,
                     int *stop_out);

/* Batch form used by the parity tests: reads r = 0..n-1 live at seq+off[r] .. seq+off[r+1]
 * (qual likewise, same offsets; qual==NULL means "no trimming", the FASTA path process_fagz
 * :849-852 which needs length > 30).  out_final[r] = final_targ or -1 (dropped). */
void kor_classify_batch(kor_sample *s, const char *seq, const char *qual, const uint64_t *off,
                        size_t n, int32_t *out_final, int32_t *out_start, int32_t *out_stop);

/* process_fqgz:762-816 over one .fastq.gz; returns 0, or -255 where the reference exit(255)s,
 * or -134 where it aborts (qual shorter than seq). */
int kor_run_fastq_gz(kor_sample *s, const char *path, FILE *reads_out);
/* main():1040-1043 */
int kor_write_result(const kor_sample *s, const char *path);

#ifdef __cplusplus
}
#endif
#endif
