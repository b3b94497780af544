#include "read_reader.hpp"
#include "gz_lines.hpp"
#include "../../include/kmer_id.h"

#include <chrono>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace kidhost {

namespace {
// Page-locking memory is slow (a fraction of a millisecond per MiB), and a run opens two readers per
// sample: batch buffers go back to this process-wide list instead of to cudaFreeHost.
std::mutex g_pinned_mu;
std::vector<std::pair<void *, size_t>> g_pinned_free;

void *pinned_get(size_t bytes)
{
    {
        std::lock_guard<std::mutex> lk(g_pinned_mu);
        for (size_t i = 0; i < g_pinned_free.size(); i++)
            if (g_pinned_free[i].second == bytes) {
                void *p = g_pinned_free[i].first;
                g_pinned_free[i] = g_pinned_free.back();
                g_pinned_free.pop_back();
                return p;
            }
    }
    void *p = nullptr;
    if (kid_host_alloc(&p, bytes) != 0) {
        fprintf(stderr, "kmer_id_b200: %s\n", kid_last_error());
        exit(1);
    }
    return p;
}

void pinned_put(void *p, size_t bytes)
{
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    g_pinned_free.emplace_back(p, bytes);
}
size_t batch_buffer_bytes(size_t max_bytes) { return max_bytes + kRefLineLimit + 16; }
} // namespace

void prewarm_batch_buffers(size_t max_bytes, int readers, bool with_quality, int depth)
{
    const size_t bytes = batch_buffer_bytes(max_bytes);
    std::vector<void *> got;
    for (int i = 0; i < readers * depth * (with_quality ? 2 : 1); i++) {
        void *p = nullptr;
        if (kid_host_alloc(&p, bytes) != 0) break; // the reader reports the failure when it needs the buffer
        got.push_back(p);
    }
    for (void *p : got) pinned_put(p, bytes);
}

ReadBatchReader::ReadBatchReader(ReadFormat fmt, const std::string &path, size_t max_reads, size_t max_bytes, int depth,
                                 unsigned gz_threads)
    : fmt_(fmt), max_reads_(max_reads), max_bytes_(max_bytes), gz_threads_(gz_threads)
{
    const bool fastq = fmt == ReadFormat::GzFastq || fmt == ReadFormat::PlainFastq;
    for (int i = 0; i < depth; i++) {
        auto b = std::make_unique<ReadBatch>();
        b->cap_bytes = max_bytes_ + kRefLineLimit;
        b->has_qual = fastq;
        b->seq = (uint8_t *)pinned_get(batch_buffer_bytes(max_bytes_));
        b->qual = fastq ? (uint8_t *)pinned_get(batch_buffer_bytes(max_bytes_)) : nullptr;
        free_.push_back(b.get());
        pool_.push_back(std::move(b));
    }
    th_ = std::thread(&ReadBatchReader::run, this, path);
}

ReadBatchReader::~ReadBatchReader()
{
    {
        std::lock_guard<std::mutex> lk(mu_);
        finished_ = true; // consumer is gone: let the producer run dry into recycled buffers
        while (!ready_.empty()) { free_.push_back(ready_.front()); ready_.pop_front(); }
    }
    cv_.notify_all();
    if (th_.joinable()) th_.join();
    for (auto &b : pool_) { pinned_put(b->seq, batch_buffer_bytes(max_bytes_)); pinned_put(b->qual, batch_buffer_bytes(max_bytes_)); }
}

ReadBatch *ReadBatchReader::get_free()
{
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return !free_.empty(); });
    ReadBatch *b = free_.front();
    free_.pop_front();
    b->n = 0;
    b->off.assign(1, 0);
    b->names.clear();
    b->name_off.assign(1, 0);
    b->last = false;
    return b;
}

void ReadBatchReader::publish(ReadBatch *b)
{
    {
        std::lock_guard<std::mutex> lk(mu_);
        if (finished_) free_.push_back(b); else ready_.push_back(b);
    }
    cv_.notify_all();
}

ReadBatch *ReadBatchReader::next()
{
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return !ready_.empty(); });
    ReadBatch *b = ready_.front();
    ready_.pop_front();
    return b;
}

void ReadBatchReader::recycle(ReadBatch *b)
{
    {
        std::lock_guard<std::mutex> lk(mu_);
        free_.push_back(b);
    }
    cv_.notify_all();
}

// one record as the reference hands it to process_qual (qual != NULL) or process_read (FASTA)
void ReadBatchReader::emit(const char *acc, size_t acclen, const char *seq, size_t seqlen, const char *qual)
{
    ReadBatch *b = cur_;
    if (seqlen > b->cap_bytes) {
        fprintf(stderr, "kmer_id_b200: a %zu-base sequence exceeds the %zu-byte batch buffer\n", seqlen, b->cap_bytes);
        exit(1);
    }
    if (b->off.back() + seqlen > b->cap_bytes || b->n >= max_reads_) {
        publish(b);
        b = cur_ = get_free();
    }
    const uint64_t o = b->off.back();
    memcpy(b->seq + o, seq, seqlen);
    if (qual) memcpy(b->qual + o, qual, seqlen);
    b->off.push_back(o + seqlen);
    b->names.insert(b->names.end(), acc, acc + acclen);
    b->name_off.push_back((uint32_t)b->names.size());
    b->n++;
    if (b->off.back() >= max_bytes_) {
        publish(b);
        cur_ = get_free();
    }
}

void ReadBatchReader::run(std::string path)
{
    cur_ = get_free();
    switch (fmt_) {
    case ReadFormat::GzFastq: run_gz_fastq(path); break;
    case ReadFormat::PlainFastq: run_plain_fastq(path); break;
    case ReadFormat::GzFasta: run_gz_fasta(path); break;
    case ReadFormat::PlainFasta: run_plain_fasta(path); break;
    }
    cur_->last = true;
    publish(cur_);
}

void ReadBatchReader::run_gz_fastq(const std::string &path)
{
    GzLineBlocks src(path, 4u << 20, gz_threads_);
    LineBlock text;
    int mod4 = 0; // :768
    const char *seq = nullptr;
    size_t seqlen = 0;
    std::string seq_carry, acc; // a record may straddle two runs of lines
    bool seq_in_carry = false;
    auto lines = [&](const char *p, const char *end) {
        while (p < end) {
            const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
            size_t len = (size_t)(eol - p);
            if (len > 0 && p[len - 1] == '\r') len--; // :786-787
            if (len > 0) {                            // :788 - empty lines do not advance mod4
                if (mod4 == 1) {
                    seq = p; seqlen = len; seq_in_carry = false;
                } else if (mod4 == 0) {
                    acc.assign(p, len);
                } else if (mod4 == 3) {
                    if (len < seqlen) // qual.at(stop) throws std::out_of_range (:729): the reference aborts
                        throw std::out_of_range("basic_string::at: quality line shorter than its read");
                    emit(acc.data(), acc.size(), seq_in_carry ? seq_carry.data() : seq, seqlen, p);
                }
                mod4 = (mod4 + 1) % 4; // :802
            }
            p = eol + 1;
        }
        if ((mod4 == 2 || mod4 == 3) && !seq_in_carry) { // the sequence line lives in this run, about to go
            seq_carry.assign(seq, seqlen);
            seq_in_carry = true;
        }
    };
    const bool stats = getenv("KID_READER_STATS") != nullptr; // where a reader thread spends its time
    double t_wait = 0, t_parse = 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    for (;;) {
        const double t0 = stats ? now() : 0;
        if (!src.next(text)) break;
        const double t1 = stats ? now() : 0;
        lines(text.head.data(), text.head.data() + text.head.size());
        lines(text.body, text.body + text.body_len);
        if (stats) { t_wait += t1 - t0; t_parse += now() - t1; }
    }
    if (stats) fprintf(stderr, "[reader] %s: waited %.3f s for inflated text, parsed/copied for %.3f s\n", path.c_str(), t_wait, t_parse);
}

// kmer_read_m3.cpp:895-931 - getline, first blank-delimited token of each line
void ReadBatchReader::run_plain_fastq(const std::string &path)
{
    std::ifstream fin(path);
    if (!fin) { open_failed_ = true; return; }
    std::string line, lseq, seq, acc;
    int mod4 = 0;
    while (std::getline(fin, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::stringstream ls(line);
        ls >> lseq;
        if (lseq.length() > 0) {
            if (mod4 == 1) seq = lseq;
            else if (mod4 == 0) acc = lseq;
            else if (mod4 == 3) {
                if (lseq.size() < seq.size())
                    throw std::out_of_range("basic_string::at: quality line shorter than its read");
                emit(acc.data(), acc.size(), seq.data(), seq.size(), lseq.data());
            }
            mod4 = (mod4 + 1) % 4;
        }
    }
}

// kmer_read_m3.cpp:780-839 - gz lines; '>' lines start a record, other lines are concatenated
void ReadBatchReader::run_gz_fasta(const std::string &path)
{
    GzLineBlocks src(path, 4u << 20, gz_threads_);
    LineBlock text;
    std::string sequence, acc;
    auto lines = [&](const char *p, const char *end) {
        while (p < end) {
            const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
            size_t len = (size_t)(eol - p);
            if (len > 0 && p[len - 1] == '\r') len--;
            if (len > 0) {
                if (p[0] == '>') {
                    if (sequence.length() > KID_KSIZE) emit(acc.data(), acc.size(), sequence.data(), sequence.size(), nullptr);
                    sequence.clear();
                    acc.assign(p + 1, len - 1);
                } else {
                    sequence.append(p, len);
                }
            }
            p = eol + 1;
        }
    };
    while (src.next(text)) {
        lines(text.head.data(), text.head.data() + text.head.size());
        lines(text.body, text.body + text.body_len);
    }
    if (sequence.length() > KID_KSIZE) emit(acc.data(), acc.size(), sequence.data(), sequence.size(), nullptr);
}

// kmer_read_m3.cpp:933-972 - getline, first blank-delimited token of each line
void ReadBatchReader::run_plain_fasta(const std::string &path)
{
    std::ifstream fin(path);
    if (!fin) { open_failed_ = true; return; }
    std::string line, lseq, sequence, acc;
    while (std::getline(fin, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::stringstream ls(line);
        ls >> lseq;
        if (!lseq.empty() && lseq[0] == '>') {
            if (sequence.length() > KID_KSIZE) emit(acc.data(), acc.size(), sequence.data(), sequence.size(), nullptr);
            sequence.clear();
            acc = lseq.substr(1);
        } else {
            sequence += lseq;
        }
    }
    if (sequence.length() > KID_KSIZE) emit(acc.data(), acc.size(), sequence.data(), sequence.size(), nullptr);
}

} // namespace kidhost
