/*
 * kmer_id.h - C-ABI of libkmerid_b200.so, the sm_100a implementation of kmer_id's
 * read-classification hot path.
 *
 * The reference (/root/reference/newkmer_10nx.cpp) has no plugin or FFI layer: it is one
 * executable with global state.  Each entry point below therefore replaces an *internal seam* of
 * that file, named per function as "replaces newkmer_10nx.cpp:<lines>".  A maintainer of the
 * reference would bind these from main()/process_fqgz().
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative KID_E* code;
 *     kid_last_error() returns the message for the calling thread's most recent failure.
 *   - "device pointer" arguments are CUDA global-memory addresses on the db's device.
 *   - there is NO CPU fallback: without a usable CUDA device every call fails with KID_ECUDA.
 *   - a kid_db is immutable after build and may be shared by several kid_sample objects; a
 *     kid_sample is used from one host thread at a time, except that different threads may drive
 *     DIFFERENT asynchronous slots of one sample at once (R1 and R2 of a sample: kid_classify_*_async
 *     and kid_wait only; begin / counts need all of them quiet).
 *   - `stream` arguments are a cudaStream_t passed as void* (NULL = the legacy default stream).
 */
#ifndef KMER_ID_H
#define KMER_ID_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KID_KSIZE 30 /* newkmer_10nx.cpp:43 */

#define KID_OK 0
#define KID_EINVAL (-1)  /* bad argument */
#define KID_ECUDA (-2)   /* CUDA runtime error / no device */
#define KID_ENOMEM (-3)  /* host or device allocation failed */
#define KID_ERANGE (-4)  /* taxon id or tree node outside [0, n_taxa) (UB in the reference) */
#define KID_ETREE (-5)   /* taxonomy has a cycle that never reaches the root (reference hangs) */
#define KID_EFULL (-6)   /* table could not place every key (reference: "out of memory in table") */
#define KID_EUNSUPPORTED (-7) /* kid_fastq_load_gz_file: a file for the host reader; nothing was counted */

/* flags for kid_db_build */
#define KID_DB_ACCEPT_U 1u /* reads: U/u count as T (kmer_read_vf6.cpp:496-500,521-525) */
/* Table layout.  Default = "M": sectors addressed by the k-mer's minimizer so that neighbouring
 * k-mers of a read share DRAM lines.  KID_DB_LAYOUT_KEYHASH = "K": one key-hashed 32-byte sector per
 * lookup (the straightforward design, kept for the ncu bake-off in profiles/). Results are identical. */
#define KID_DB_LAYOUT_KEYHASH 2u

typedef struct kid_db kid_db;         /* GPU-resident probe table + taxonomy tree */
typedef struct kid_sample kid_sample; /* per-sample accumulators: gcount, seen flags, counters */

const char *kid_last_error(void);
int kid_device_count(int *n);
/* Creates the CUDA context of `device` (seconds on a large box).  Optional: kid_db_build does it on
 * first use; a host calls this from a helper thread so that it overlaps reading the probe file
 * (the reference spends that time in its 24 GiB table memset, newkmer_10nx.cpp:943,186-189). */
int kid_device_init(int device);
const char *kid_version(void);
/* number of CUDA kernels this library has launched in this process (monotonic) */
unsigned long long kid_kernel_launches(void);
/* page-locked host memory for the batch buffers handed to kid_classify_host (replaces the
 * per-line std::string of process_fqgz, newkmer_10nx.cpp:785).  Page-locked through the context of
 * the device last named to kid_device_init / kid_db_build (device 0 before either), usable from all. */
int kid_host_alloc(void **p, size_t bytes);
void kid_host_free(void *p);

/* ---- database ------------------------------------------------------------------------------
 * kid_db_build: replaces Hashtable::Hashtable/add_kmer (newkmer_10nx.cpp:173-180, 235-263) as
 * driven by process_kmer (:619-661), and Tree1 (:93-154) as filled by main():973-983.
 *   keys[i], taxa[i]  i = 0..n_keys-1, in FILE ORDER: one entry per forward 30-window of every
 *                     parsed probe line, key = the 60-bit forward encoding (:634-646).
 *                     Semantics reproduced: the first entry of a key wins, entries with
 *                     taxa[i] == 0 are invisible.  taxa[i] >= n_taxa -> KID_ERANGE.
 *   parent[t]         t = 0..n_taxa-1: Tree1::parent after all add_edge calls (default 1, :103).
 *   keys_on_device    non-zero if keys/taxa are device pointers (then they are read in place).
 *   log2_sectors      0 = size the table from n_keys; else log2 of the number of 32-byte sectors
 *                     (layout M: 12..32, layout K: 22..32).  The default is a function of n_keys
 *                     alone (<= 0.225 keys per 3-entry sector: 16 GiB for 1.09e8 keys; KID_DB_DENSE=1
 *                     halves that for ~3 % fewer lookups/s); it shrinks, with a message on stderr,
 *                     only when the device cannot hold it.
 *   n_keys            at most 2^32-2 entries (owner indices are 32-bit during the build).
 * Environment (tuning/test knobs): KID_DB_SUB_BITS=2|3|4 forces how many sectors (4, 8, 16) one
 * minimizer addresses in layout M; by default 4, and 8 or 16 only for databases beyond 4e8 keys
 * that device memory keeps densely packed (kid_table2.cuh).  KID_DB_MM=20 addresses the table by
 * 20-mer minimizers instead of 16-mers.  KID_DB_DENSE=1: see log2_sectors.
 */
int kid_db_build(const uint64_t *keys, const uint32_t *taxa, size_t n_keys, int keys_on_device,
                 const int32_t *parent, int n_taxa, int device, unsigned flags, int log2_sectors,
                 void *stream, kid_db **out);
void kid_db_free(kid_db *db);
int kid_db_n_taxa(const kid_db *db);
int kid_db_device(const kid_db *db);
/* distinct visible keys, table sectors (32 bytes each), table bytes, keys living outside their
 * home sector */
int kid_db_stats(const kid_db *db, uint64_t *n_distinct, uint64_t *n_sectors,
                 uint64_t *table_bytes, uint64_t *n_displaced);
/* Hashtable::getHash (:204-233) for n host keys -> host taxa (0 = absent). Test/diagnostic hook. */
int kid_db_lookup(const kid_db *db, const uint64_t *keys, size_t n, uint32_t *taxa_out);
/* Tree1::msca (:118-144) for n host pairs (x[i], y[i]) evaluated by the device routine. */
int kid_db_msca(const kid_db *db, const int32_t *x, const int32_t *y, size_t n, int32_t *out);
/* device address and number of 32-byte sectors of the table (for microbenchmarks such as the
 * random sector-gather ceiling in bench.py) */
int kid_db_table_device(const kid_db *db, void **table, uint64_t *n_sectors);

/* ---- per-sample state ----------------------------------------------------------------------
 * replaces the globals gcount[], ucount[], kmer_seen (newkmer_10nx.cpp:61-64) and their reset in
 * main():1017-1023.  seen flags are one bit per table slot, n_seen_words 32-bit words. */
int kid_sample_create(const kid_db *db, kid_sample **out);
void kid_sample_free(kid_sample *s);
int kid_sample_begin(kid_sample *s, void *stream); /* zero gcount/seen/counters, asynchronously */
const kid_db *kid_sample_db(const kid_sample *s);   /* the database the sample was created for */

/* ---- the hot path ---------------------------------------------------------------------------
 * kid_classify_device: replaces process_qual (:714-760) + process_read (:452-617) +
 * getHash/msca for a whole batch, ON THE GIVEN STREAM, with every buffer already on the device.
 *   seq, qual   device byte buffers; read r occupies [off[r], off[r+1]) in both (same offsets).
 *               qual == NULL: no trimming, reads need length > 30 (process_fagz :849-852).
 *               Both buffers must be readable for 16 bytes past off[n_reads] (padding) and seq
 *               must be 16-byte aligned (the kernel loads aligned 128-bit words; KID_EINVAL otherwise).
 *   off         device uint64[n_reads+1]
 *   out_taxon   device int32[n_reads] or NULL: final_targ (:616), or -1 if the read was dropped
 *               by the length rule (:755) and therefore not counted anywhere.
 *   out_span    device uint32[2*n_reads] or NULL: trimmed (start, stop) (:724-753)
 * Accumulates gcount, seen flags and the lookup/hit counters into s.  The work (kid_pack_kernel, then
 * the scan) is asynchronous on `stream`; the call itself waits once for the stream to hand back
 * off[n_reads], which sizes the packed scratch buffer. */
int kid_classify_device(kid_sample *s, const uint8_t *seq, const uint8_t *qual,
                        const uint64_t *off, size_t n_reads, int32_t *out_taxon,
                        uint32_t *out_span, void *stream);

/* kid_classify_host: same, from HOST buffers (pinned or pageable; pinned overlaps copies with
 * compute).  Splits the batch into chunks, double-buffers H2D copy / kernel / D2H copy on two
 * internal streams and returns when out_* are complete.  This is the call a FASTQ reader makes
 * per batch (replaces the per-record process_qual call at newkmer_10nx.cpp:800). */
int kid_classify_host(kid_sample *s, const uint8_t *seq, const uint8_t *qual,
                      const uint64_t *off, size_t n_reads, int32_t *out_taxon,
                      uint32_t *out_span);
/* tuning knob for kid_classify_host: reads per chunk (default 1<<18) */
int kid_sample_set_chunk_reads(kid_sample *s, size_t chunk_reads);
/* bytes moved by kid_classify_host since kid_sample_begin (for bench.py's e2e accounting) */
int kid_sample_transfer_bytes(const kid_sample *s, uint64_t *h2d, uint64_t *d2h);

/* ---- packed read batches ----------------------------------------------------------------------
 * What the k-mer scan actually consumes, and what a host parser should ship instead of text: per read
 * only the bases that survive process_qual's trim (:714-760), as 2-bit codes.  150-base reads cost
 * 48 bytes each on the wire instead of 308 (bases + qualities + offset), and qualities never leave
 * the host.
 *   words   uint32 stream.  A read's trimmed bases seq[start..stop], 16 per word, first base in the
 *           top bit pair, A/a 0, C/c 1, G/g 2, T/t 3 (U/u 3 with KID_DB_ACCEPT_U), any other byte 0;
 *           the last word zero padded.  If (and only if) the read has KID_PK_INVALID set,
 *           ceil(tlen/32) validity words follow its code words: first base in the top bit,
 *           1 = the byte was one of the accepted letters (:480-524), zero padded.
 *   meta    uint32[2*(n_reads+1)]: meta[2r] = index of the read's first word | KID_PK_INVALID,
 *           meta[2r+1] = tlen = stop-start+1, or 0 for a read the length rule drops
 *           (stop-start < 30, :755 - such a read needs no words).  Entry n_reads = {end index, 0}.
 *           Word indices are non-decreasing and below 2^31; gaps between reads are allowed.
 * kid_pack_reads writes this format on the host (the trim is process_qual's, on the calling thread,
 * the packing is SIMD); kid_classify_device / kid_classify_host write it on the device with
 * kid_pack_kernel when they are given text. */
#define KID_PK_INVALID 0x80000000u
#define KID_PACK_IMPL_BYTES 0x100u /* kid_pack_reads: force the byte-loop packer (the format's definition) */
#define KID_PACK_IMPL_SWAR 0x200u  /* ... the 64-bit SWAR packer instead of AVX2 */
/* words that n_reads reads totalling `bases` bases can need at most */
size_t kid_pack_bound(size_t n_reads, uint64_t bases);
/* Host-side packer: replaces process_qual (:714-760) and the base switch of process_read
 * (:477-525) for reads [0, n_reads) of a text batch laid out as for kid_classify_host.
 *   flags      KID_DB_ACCEPT_U or 0 (| KID_PACK_IMPL_* in tests: all implementations agree)
 *   word0      index the first word written gets in meta (the caller appends batch after batch)
 *   words      receives the words; words_cap entries available (see kid_pack_bound)
 *   meta       2*(n_reads+1) entries
 *   span       2*n_reads entries or NULL: (start, stop) as process_qual leaves them
 *   n_words    out: words written
 * Reads and writes host memory only; thread safe. */
int kid_pack_reads(const uint8_t *seq, const uint8_t *qual, const uint64_t *off, size_t n_reads,
                   unsigned flags, uint32_t word0, uint32_t *words, size_t words_cap, uint32_t *meta,
                   uint32_t *span, size_t *n_words);
/* Device-side packer (kid_pack_kernel): the same conversion for a text batch that is already on the
 * device (buffers as for kid_classify_device, off[0] = 0, total_bases = off[n_reads]).  Read r's words
 * start at off[r]/16 + off[r]/32 + 2r, so words_cap must be at least kid_pack_bound(n_reads,
 * total_bases); the gaps are never read.  Long reads (> 496 bases) always carry validity words.
 * Asynchronous on `stream`. */
int kid_pack_device(const kid_db *db, const uint8_t *seq, const uint8_t *qual, const uint64_t *off,
                    uint64_t total_bases, size_t n_reads, uint32_t *words, size_t words_cap,
                    uint32_t *meta, uint32_t *out_span, void *stream);
/* kid_classify_packed_device: process_read (:452-617) + getHash/msca for a packed batch whose words
 * and meta are ALREADY on the device, on the given stream.  meta[].x word indices are relative to
 * `words`.  out_taxon: device int32[n_reads] or NULL, -1 for dropped reads.  Asynchronous. */
int kid_classify_packed_device(kid_sample *s, const uint32_t *words, const uint32_t *meta,
                               size_t n_reads, int32_t *out_taxon, void *stream);
/* kid_classify_packed_host: same from HOST buffers, chunked and double-buffered like
 * kid_classify_host; returns when out_taxon (host int32[n_reads] or NULL) is complete.  words[0] is
 * the word with index word0 (= the word0 given to kid_pack_reads). */
int kid_classify_packed_host(kid_sample *s, const uint32_t *words, uint32_t word0, const uint32_t *meta,
                             size_t n_reads, int32_t *out_taxon);

/* ---- dense read batches: the fewest bytes on the wire ------------------------------------------
 * 41.5 bytes per 150-base read instead of 50.6: no padding at all, one 32-bit offset per read, and the
 * rare non-ACGT bases as a list of positions instead of validity words.
 *   codes     uint32 stream of 2-bit codes, 16 per word, first base in the top bit pair; read r occupies
 *             bases [boff[r], boff[r+1]) of the stream (its trimmed span; a dropped read has none).
 *             codes[0] holds bases 16*(base0/16) .. of the stream (base0 = what kid_pack_reads_dense got).
 *   boff      uint32[n_reads+1]
 *   flagbits  bit (read0 + r) % 32 of word (read0 + r) / 32: read r contains a base outside ACGT
 *             (kid_pack_reads_dense zeroes a word when it packs the word's first read)
 *   inv       uint32[n_inv]: stream positions of those bases, ascending
 * kid_pack_reads_dense appends n_reads reads to a batch (base0 / read0 = bases / reads already in it; the
 * arrays it is given start at the batch's beginning for flagbits, at the append position for the
 * rest).  The device turns a dense batch into a packed one with one streaming kernel
 * (kid_expand_kernel) before the scan.  Limits: < 2^32 bases per batch. */
size_t kid_dense_bound(uint64_t bases); /* words of `codes` that `bases` bases need at most */
int kid_pack_reads_dense(const uint8_t *seq, const uint8_t *qual, const uint64_t *off, size_t n_reads,
                         unsigned flags, uint32_t base0, uint32_t *codes, size_t codes_cap, uint32_t *boff,
                         uint32_t *flagbits, size_t read0, uint32_t *inv, size_t inv_cap, size_t *n_inv,
                         uint32_t *span, uint32_t *n_bases);
/* Host buffers, chunked over the slots, returns when out_taxon (host int32[n_reads] or NULL) is complete.
 * kid_sample_set_chunk_reads values are rounded to multiples of 32 reads here. */
int kid_classify_dense_host(kid_sample *s, const uint32_t *codes, const uint32_t *boff, const uint32_t *flagbits,
                            const uint32_t *inv, size_t n_inv, size_t n_reads, int32_t *out_taxon);
int kid_classify_dense_async(kid_sample *s, int slot, const uint32_t *codes, const uint32_t *boff,
                             const uint32_t *flagbits, const uint32_t *inv, size_t n_inv, size_t n_reads,
                             int32_t *out_taxon);

/* ---- asynchronous slots -------------------------------------------------------------------------
 * One host thread pipelines parse | H2D | kernels | D2H: it fills a pinned buffer, submits it on a
 * slot and goes on parsing; kid_wait(slot) returns when that slot's outputs are complete and its
 * input buffers may be reused.  A slot is one CUDA stream plus device staging buffers owned by the
 * sample; submissions on one slot run in order, different slots overlap.  The caller owns the host
 * buffers (kid_host_alloc) and must not touch them between submit and kid_wait.  Replaces the
 * synchronous per-record call at newkmer_10nx.cpp:798-801. */
#define KID_MAX_SLOTS 4
int kid_classify_async(kid_sample *s, int slot, const uint8_t *seq, const uint8_t *qual,
                       const uint64_t *off, size_t n_reads, int32_t *out_taxon, uint32_t *out_span);
int kid_classify_packed_async(kid_sample *s, int slot, const uint32_t *words, uint32_t word0,
                              const uint32_t *meta, size_t n_reads, int32_t *out_taxon);
int kid_wait(kid_sample *s, int slot);

/* ---- sample end -------------------------------------------------------------------------------
 * kid_sample_counts: replaces the read-out loop main():1040-1043.  Computes ucount as the
 * per-taxon histogram of seen flags (equivalent to :596-603) and copies
 * gcount/ucount (int32[n_taxa]) to the host.  Synchronous.  Either pointer may be NULL. */
int kid_sample_counts(kid_sample *s, int32_t *gcount, int32_t *ucount, void *stream);
/* Same for ONE sample whose reads were split over n kid_sample objects of the same kid_db on the
 * same device (e.g. R1 and R2 classified concurrently by two host threads): gcount is summed,
 * ucount is the histogram of the OR of their seen flags (a k-mer hit in both shards counts once). */
int kid_samples_counts(kid_sample *const *samples, int n, int32_t *gcount, int32_t *ucount, void *stream);
/* number of getHash calls (:529), hits (target > 0) and reads counted in gcount ("tct", :614) */
int kid_sample_counters(kid_sample *s, uint64_t *lookups, uint64_t *hits, uint64_t *reads,
                        void *stream);

/* ---- multi-GPU building blocks (one process per GPU; the exchange itself is the caller's:
 * NCCL / peer memory).  gcount is additive across shards; ucount is NOT: the
 * seen flags must be OR-ed across ranks before they are histogrammed. */
/* device address of gcount (int32[n_taxa]) - to be sum-all-reduced in place by the caller */
int kid_sample_gcount_device(kid_sample *s, int32_t **gcount);
/* device address of the seen bitmap (uint32[n_words]); words are padded so n_words % 1024 == 0 */
int kid_sample_seen_device(kid_sample *s, uint32_t **seen, uint64_t *n_words);
/* OR `n_src` bitmaps' word range [word0, word0+n) into dst (device pointers, possibly peer
 * memory mapped over NVLink): dst[i] = src[0][word0+i] | ... ; dst may alias a src range. */
int kid_seen_or_device(const kid_db *db, uint32_t *dst, const uint32_t *const *src, int n_src,
                       uint64_t word0, uint64_t n_words, void *stream);
/* Fused form of the two calls above for peer memory: ucount_partial[t] += #slots of taxon t whose
 * flag is set in ANY of the n_src bitmaps, for words [word0, word0+n).  The sources are read in
 * place - with NVLink-mapped peer pointers (CUDA IPC / symmetric memory) the OR-reduction and the
 * histogram are one kernel and no bitmap is copied or written.  n_src <= 16. */
int kid_ucount_or_range_device(const kid_db *db, const uint32_t *const *seen_srcs, int n_src,
                               uint64_t word0, uint64_t n_words, int32_t *ucount_partial, void *stream);
/* ---- one host process, several GPUs (the C++ hosts; no torch, no IPC) ---------------------------
 * Reads of ONE sample may be dealt to kid_sample objects on different GPUs, each over its own replica
 * of the same database (kid_db_build is deterministic: replicas agree slot for slot).  Sample end:
 *   1. kid_device_sync every GPU (all seen bits written);
 *   2. on GPU r: kid_sample_ucount_partial(shard[r], shards, n, r, n) - ONE kernel that reads word
 *      range r of every shard's seen bitmap IN PLACE over NVLink (peer access, kid_peer_enable), ORs
 *      them and histograms the result into that shard's ucount buffer (zeroed first);
 *   3. sum gcount and the partial ucounts over the GPUs: ncclAllReduce in place on the device buffers
 *      (kid_sample_gcount_device / kid_sample_ucount_device), or read them back and add;
 *   4. kid_sample_read_counts on any shard.
 * kmer_id_b200/host/multi_gpu.hpp does exactly this. */
/* peer access between every ordered pair of the listed devices (idempotent) */
int kid_peer_enable(const int *devices, int n);
int kid_sample_ucount_partial(kid_sample *s, kid_sample *const *shards, int n_shards, int part, int n_parts,
                              void *stream);
int kid_sample_ucount_device(kid_sample *s, int32_t **ucount);
/* plain device -> host copies of the sample's gcount / ucount buffers as they are (no histogram pass) */
int kid_sample_read_counts(kid_sample *s, int32_t *gcount, int32_t *ucount, void *stream);
int kid_device_sync(int device);
/* Make the sample keep its seen flags in caller-owned device memory of at least n_words 32-bit
 * words (e.g. a symmetric-memory allocation that peers can map).  The buffer is cleared by
 * kid_sample_begin like the internal one and is not freed by the library. */
int kid_sample_use_seen_buffer(kid_sample *s, uint32_t *buf, uint64_t n_words);
/* ucount_partial[t] += #set flags in seen[word0 .. word0+n) whose slot holds taxon t
 * (device int32[n_taxa], NOT zeroed here).  With disjoint word ranges per rank the partial
 * histograms are additive, so a sum-all-reduce finishes the job. */
int kid_ucount_range_device(const kid_db *db, const uint32_t *seen, uint64_t word0,
                            uint64_t n_words, int32_t *ucount_partial, void *stream);

/* ---- gzip FASTQ files read on the device -----------------------------------------------------------
 * Replaces process_fqgz (newkmer_10nx.cpp:762-816: gzread, line splitting, the 4-line state) together
 * with process_qual/process_read for a WHOLE file: the compressed bytes are copied to the device and
 * inflated there (csrc/kid_inflate.cuh: speculative inflate of 32 KiB pieces to 16-bit symbols, chained
 * and CRC-checked), lines are framed with two prefix sums (newlines; non-empty lines, whose count mod 4
 * is the reference's mod4), and the records go through kid_pack_kernel and the k-mer scan.  The host
 * only sees per-read taxa and, on request, the header and trimmed bases of chosen reads.
 *   kid_fastq_load_gz_file  KID_OK: the file is on the device, framed; *n_reads = its records.
 *                           KID_EUNSUPPORTED: a file this path does not reproduce byte for byte (no gzip
 *                           header, fixed/stored-only deflate streams, trailing bytes or a bad CRC-32 /
 *                           ISIZE - zlib has to judge those -, a line of >= 16 KiB (fatal at :773), a
 *                           quality line shorter than its read (the reference aborts at :729), more text
 *                           than device memory holds: the whole file is inflated at once, ~45 bytes of
 *                           device memory per compressed byte).  Nothing has been counted: read
 *                           the file with the host reader, whose error behaviour is the reference's.
 *   kid_fastq_prefetch_gz_file  optional: starts reading `path` into a second device buffer on a helper thread
 *                           and returns at once; a later kid_fastq_load_gz_file of the same path finds the
 *                           bytes there (the host reads sample i+1 while the device works on sample i).
 *   kid_fastq_classify      all records of the loaded file into sample s (same device); out_taxon: host
 *                           int32[n_reads] (page-locked for speed) or NULL, -1 = dropped by the length rule.
 *   kid_fastq_fetch         for n record indices: *lens = uint32[2n] (header length, trimmed bases
 *                           length), *data = header_0 bases_0 header_1 bases_1 ... (no separators; the
 *                           header is the whole '@' line without its line end, what :796 keeps as acc);
 *                           both live in the object until its next call.  For _reads.txt (:608-611).
 *   kid_fastq_stats         text bytes, pieces, pieces inflated twice, gzip members, and the seconds of the
 *                           last file's phases: read+copy, find, inflate, chain, resolve+CRC, frame,
 *                           classify, fetch (n_phases <= 8).
 * One object serves one host thread; objects on different streams overlap (R1 and R2 of a sample).
 * Environment: KID_GZ_GPU_PIECE (bytes, 32768), KID_GZ_GPU_EXPAND (symbols reserved per compressed byte, 8). */
typedef struct kid_fastq kid_fastq;
int kid_fastq_create(const kid_db *db, kid_fastq **out);
void kid_fastq_free(kid_fastq *f);
int kid_fastq_prefetch_gz_file(kid_fastq *f, const char *path);
int kid_fastq_load_gz_file(kid_fastq *f, const char *path, size_t *n_reads);
int kid_fastq_classify(kid_fastq *f, kid_sample *s, int32_t *out_taxon);
int kid_fastq_fetch(kid_fastq *f, const uint32_t *reads, size_t n, const char **data, const uint32_t **lens);
int kid_fastq_stats(const kid_fastq *f, uint64_t *n_text, uint64_t *n_pieces, uint64_t *n_again, uint64_t *n_members,
                    double *phase_seconds, int n_phases);

#ifdef __cplusplus
}
#endif
#endif /* KMER_ID_H */
