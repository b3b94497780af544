// fastq_reader.hpp - gzip FASTQ -> batches of reads laid out for kid_classify_host.
// Replaces process_fqgz (newkmer_10nx.cpp:762-816) up to the point where it calls process_qual.
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

// Pinned host buffers sized once; reads are concatenated with no separators.
struct ReadBatch {
    uint8_t *seq = nullptr;   // pinned, cap_bytes + 16
    uint8_t *qual = nullptr;  // pinned, same offsets as seq (only the first seqlen bytes of a
                              // quality line are ever looked at, :724-753)
    std::vector<uint64_t> off;      // n + 1
    std::vector<char> names;        // header lines (with '@'), concatenated
    std::vector<uint32_t> name_off; // n + 1
    size_t n = 0;
    size_t cap_bytes = 0;
    bool last = false;        // no more batches after this one
};

class FastqBatchReader {
public:
    // starts a background thread that inflates + parses `path` into batches of at most
    // max_reads reads / max_bytes bases, keeping at most `depth` batches ahead of the consumer
    FastqBatchReader(const std::string &path, size_t max_reads, size_t max_bytes, int depth = 3);
    ~FastqBatchReader();
    // blocks until the next batch is ready; returns nullptr after the last one was handed out
    ReadBatch *next();
    void recycle(ReadBatch *b);

private:
    void run(std::string path);
    ReadBatch *get_free();
    void publish(ReadBatch *b);

    size_t max_reads_, max_bytes_;
    std::vector<std::unique_ptr<ReadBatch>> pool_;
    std::deque<ReadBatch *> free_, ready_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::thread th_;
    bool finished_ = false;
};

} // namespace kidhost
