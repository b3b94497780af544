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
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace kidhost {

enum class ReadFormat { GzFastq, PlainFastq, GzFasta, PlainFasta };

// Pinned host buffers sized once; reads are concatenated with no separators.
struct ReadBatch {
    uint8_t *seq = nullptr;   // pinned, cap_bytes + 16
    uint8_t *qual = nullptr;  // pinned, same offsets as seq (only the first seqlen bytes of a
                              // quality line are ever looked at, :724-753); unused for FASTA
    std::vector<uint64_t> off;      // n + 1
    std::vector<char> names;        // header lines as the reference keeps them, concatenated
    std::vector<uint32_t> name_off; // n + 1
    size_t n = 0;
    size_t cap_bytes = 0;
    bool has_qual = true;
    bool last = false;        // no more batches after this one
};

// Page-locks the buffers that `readers` ReadBatchReaders of this max_bytes/depth will ask for and
// parks them in the process-wide pool (read_reader.cpp); meant for a helper thread at start-up.
void prewarm_batch_buffers(size_t max_bytes, int readers, bool with_quality, int depth = 3);

class ReadBatchReader {
public:
    // starts a background thread that inflates/parses `path` into batches of at most max_reads
    // reads / max_bytes bases, keeping at most `depth` batches ahead of the consumer
    // gz_threads: inflate workers for gz inputs (0 = default_gz_threads() of pgz.hpp)
    ReadBatchReader(ReadFormat fmt, const std::string &path, size_t max_reads, size_t max_bytes, int depth = 3,
                    unsigned gz_threads = 0);
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
    ReadBatch *get_free();
    void publish(ReadBatch *b);

    ReadFormat fmt_;
    size_t max_reads_, max_bytes_;
    unsigned gz_threads_ = 0;
    std::vector<std::unique_ptr<ReadBatch>> pool_;
    std::deque<ReadBatch *> free_, ready_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::thread th_;
    ReadBatch *cur_ = nullptr;
    bool finished_ = false;
    bool open_failed_ = false;
};

} // namespace kidhost
