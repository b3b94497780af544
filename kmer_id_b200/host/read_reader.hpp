// read_reader.hpp - read files -> batches laid out for kid_classify_host.
// Replaces the record loops of the reference readers up to the point where they call
// process_qual / process_read:
//   GzFastq    process_fqgz  newkmer_10nx.cpp:762-816 (same in kmer_read_m3.cpp:724-778, kmer_read_vf6.cpp)
//   PlainFastq process_fq    kmer_read_m3.cpp:895-931
//   GzFasta    process_fagz  kmer_read_m3.cpp:780-839   (newkmer_10nx.cpp:818-875)
//   PlainFasta process_fa    kmer_read_m3.cpp:933-972   (newkmer_10nx.cpp:877-913)
#pragma once
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace kidhost {

enum class ReadFormat { GzFastq, PlainFastq, GzFasta, PlainFasta };

// What a batch carries for the GPU.
//   Packed (the hosts): every record is trimmed and 2-bit packed by kid_pack_reads_dense while its lines
//           are still in cache; codes/boff/flagbits/inv/taxon are pinned and go to
//           kid_classify_dense_async as they are (41.5 bytes per 150-base read).  Qualities are never
//           copied; the bases are kept in ordinary memory for _reads.txt.
//   Text    (tests, and callers of kid_classify_host): bases and qualities in pinned buffers.
enum class BatchMode { Packed, Text };

// Buffers sized once; reads are concatenated with no separators.
struct ReadBatch {
    uint8_t *seq = nullptr;   // cap_bytes + 16 (pinned in Text mode)
    uint8_t *qual = nullptr;  // Text mode, FASTQ: pinned, same offsets as seq (only the first seqlen
                              // bytes of a quality line are ever looked at, :724-753)
    // Packed mode: the dense batch of include/kmer_id.h, all pinned
    uint32_t *codes = nullptr;    // 2-bit stream, n_bases used of cap_bytes
    uint32_t *boff = nullptr;     // n + 1 stream positions
    uint32_t *flagbits = nullptr; // one bit per read
    uint32_t *inv = nullptr;      // n_inv positions of non-ACGT bases, cap_inv
    int32_t *taxon = nullptr;     // cap_reads: where the GPU's per-read result lands
    std::vector<uint32_t> span;     // Packed mode: 2n, (start, stop) as process_qual leaves them
    std::vector<uint64_t> off;      // n + 1
    std::vector<char> names;        // header lines as the reference keeps them, concatenated
    std::vector<uint32_t> name_off; // n + 1
    size_t n = 0, n_inv = 0;
    uint32_t n_bases = 0;
    size_t cap_bytes = 0, cap_inv = 0, cap_reads = 0;
    bool has_qual = true;
    bool last = false;        // no more batches after this one
    int slot = -1;            // for the consumer: the asynchronous slot this batch was submitted on
};

// Page-locks the buffers that `readers` ReadBatchReaders of this shape will ask for and parks them in
// the process-wide pool (read_reader.cpp); meant for a helper thread at start-up.
void prewarm_batch_buffers(size_t max_reads, size_t max_bytes, int readers, int depth = 4);

// batches a reader needs so that `slots` submissions can be in flight while its parse workers stay busy
int pipeline_depth(int slots);

class ReadBatchReader {
public:
    // starts a background thread that inflates/parses `path` into batches of at most max_reads
    // reads / max_bytes bases, keeping at most `depth` batches ahead of the consumer
    // gz_threads: inflate workers for gz inputs (0 = default_gz_threads() of pgz.hpp).  Packed gz FASTQ
    // is parsed, trimmed and packed by several threads (KID_PARSE_THREADS, default 3; 0 = on the reader
    // thread itself): `depth` must then cover them (pipeline_depth()).
    // pack_flags: KID_DB_ACCEPT_U or 0 (Packed mode)
    ReadBatchReader(ReadFormat fmt, const std::string &path, size_t max_reads, size_t max_bytes, int depth = 3,
                    unsigned gz_threads = 0, BatchMode mode = BatchMode::Packed, unsigned pack_flags = 0);
    ~ReadBatchReader();
    ReadBatch *next();          // blocks; the batch flagged `last` ends the stream
    void recycle(ReadBatch *b);
    bool open_failed() const { return open_failed_; } // plain-text formats: the file did not open

private:
    void run(std::string path);
    void run_gz_fastq(const std::string &path);
    void run_plain_fastq(const std::string &path);
    void run_gz_fasta(const std::string &path);
    void run_plain_fasta(const std::string &path);
    void emit(const char *acc, size_t acclen, const char *seq, size_t seqlen, const char *qual);
    void emit_into(ReadBatch &b, const char *acc, size_t acclen, const char *seq, size_t seqlen, const char *qual);
    ReadBatch *get_free();
    void publish(ReadBatch *b);
    void alloc_batch(ReadBatch &b, size_t cap_bytes, size_t cap_reads);
    void free_batch(ReadBatch &b);
    void pack_record(ReadBatch &b, const char *seq, size_t seqlen, const char *qual);
    void run_gz_fastq_parallel(const std::string &path);
    void publish_ordered(uint64_t index, ReadBatch *b);

    ReadFormat fmt_;
    size_t max_reads_, max_bytes_;
    unsigned gz_threads_ = 0;
    BatchMode mode_ = BatchMode::Packed;
    unsigned pack_flags_ = 0;
    std::vector<std::unique_ptr<ReadBatch>> pool_;
    std::deque<ReadBatch *> free_, ready_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::thread th_;
    ReadBatch *cur_ = nullptr;
    std::map<uint64_t, ReadBatch *> done_; // parallel gz FASTQ: finished batches waiting for their turn
    uint64_t next_pub_ = 0;
    unsigned parse_threads_ = 0;
    bool finished_ = false;
    bool open_failed_ = false;
};

} // namespace kidhost
