#include "fastq_reader.hpp"
#include "gz_lines.hpp"
#include "../../include/kmer_id.h"

#include <stdexcept>

namespace kidhost {

FastqBatchReader::FastqBatchReader(const std::string &path, size_t max_reads, size_t max_bytes, int depth)
    : max_reads_(max_reads), max_bytes_(max_bytes)
{
    for (int i = 0; i < depth; i++) {
        auto b = std::make_unique<ReadBatch>();
        b->cap_bytes = max_bytes_ + kRefLineLimit;
        void *p = nullptr, *q = nullptr;
        if (kid_host_alloc(&p, b->cap_bytes + 16) != 0 || kid_host_alloc(&q, b->cap_bytes + 16) != 0) {
            fprintf(stderr, "nk10: %s\n", kid_last_error());
            exit(1);
        }
        b->seq = (uint8_t *)p;
        b->qual = (uint8_t *)q;
        free_.push_back(b.get());
        pool_.push_back(std::move(b));
    }
    th_ = std::thread(&FastqBatchReader::run, this, path);
}

FastqBatchReader::~FastqBatchReader()
{
    {
        std::lock_guard<std::mutex> lk(mu_);
        finished_ = true; // consumer is gone: let the producer run dry into recycled buffers
        while (!ready_.empty()) { free_.push_back(ready_.front()); ready_.pop_front(); }
    }
    cv_.notify_all();
    if (th_.joinable()) th_.join();
    for (auto &b : pool_) { kid_host_free(b->seq); kid_host_free(b->qual); }
}

ReadBatch *FastqBatchReader::get_free()
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

void FastqBatchReader::publish(ReadBatch *b)
{
    {
        std::lock_guard<std::mutex> lk(mu_);
        if (finished_) free_.push_back(b); else ready_.push_back(b);
    }
    cv_.notify_all();
}

ReadBatch *FastqBatchReader::next()
{
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return !ready_.empty(); });
    ReadBatch *b = ready_.front();
    ready_.pop_front();
    return b;
}

void FastqBatchReader::recycle(ReadBatch *b)
{
    {
        std::lock_guard<std::mutex> lk(mu_);
        free_.push_back(b);
    }
    cv_.notify_all();
}

void FastqBatchReader::run(std::string path)
{
    GzLineBlocks src(path);
    std::vector<char> text;
    ReadBatch *b = get_free();
    int mod4 = 0; // :768
    const char *seq = nullptr;
    size_t seqlen = 0;
    std::string seq_carry, acc; // a record may straddle two text blocks
    bool seq_in_carry = false;
    while (src.next(text)) {
        const char *p = text.data(), *end = p + text.size();
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
                    const char *s = seq_in_carry ? seq_carry.data() : seq;
                    if (len < seqlen) // qual.at(stop) throws std::out_of_range (:729): the reference aborts
                        throw std::out_of_range("basic_string::at: quality line shorter than its read");
                    if (b->off.back() + seqlen > b->cap_bytes || b->n >= max_reads_) {
                        publish(b);
                        b = get_free();
                    }
                    const uint64_t o = b->off.back();
                    memcpy(b->seq + o, s, seqlen);
                    memcpy(b->qual + o, p, seqlen);
                    b->off.push_back(o + seqlen);
                    b->names.insert(b->names.end(), acc.begin(), acc.end());
                    b->name_off.push_back((uint32_t)b->names.size());
                    b->n++;
                    if (b->off.back() >= max_bytes_) {
                        publish(b);
                        b = get_free();
                    }
                }
                mod4 = (mod4 + 1) % 4; // :802
            }
            p = eol + 1;
        }
        if (mod4 == 2 || mod4 == 3) { // the sequence line lives in `text`, which is about to go
            if (!seq_in_carry) { seq_carry.assign(seq, seqlen); seq_in_carry = true; }
        }
    }
    b->last = true;
    publish(b);
}

} // namespace kidhost
