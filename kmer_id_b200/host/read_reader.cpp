#include "read_reader.hpp"
#include "gz_lines.hpp"
#include "../../include/kmer_id.h"

#include <algorithm>
#include <chrono>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
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
size_t batch_code_words(size_t cap_bytes) { return kid_dense_bound(cap_bytes) + 16; }
size_t batch_inv_cap(size_t cap_bytes) { return cap_bytes / 16 + 4096; } // non-ACGT bases are rare; pack_record grows the list

// text bytes per super-block of the parallel gz FASTQ path
size_t parallel_block_bytes(size_t max_bytes) { return std::max<size_t>(4096, std::min<size_t>(max_bytes, (size_t)8 << 20)); }

unsigned parse_threads_default()
{
    if (const char *e = getenv("KID_PARSE_THREADS")) return (unsigned)std::max(0, atoi(e));
    return std::thread::hardware_concurrency() >= 8 ? 3u : (std::thread::hardware_concurrency() >= 4 ? 2u : 0u);
}
} // namespace

int pipeline_depth(int slots) { return slots + 3 + (int)parse_threads_default(); }

void prewarm_batch_buffers(size_t max_reads, size_t max_bytes, int readers, int depth)
{
    const size_t cap = max_bytes + kRefLineLimit;
    std::vector<std::pair<void *, size_t>> got;
    for (int i = 0; i < readers * depth; i++)
        for (size_t bytes : { 4 * batch_code_words(cap), 4 * (max_reads + 1), 4 * (max_reads / 32 + 2), 4 * batch_inv_cap(cap), 4 * max_reads }) {
            void *p = nullptr;
            if (kid_host_alloc(&p, bytes) != 0) break; // the reader reports the failure when it needs the buffer
            got.emplace_back(p, bytes);
        }
    for (auto &g : got) pinned_put(g.first, g.second);
}

void ReadBatchReader::alloc_batch(ReadBatch &b, size_t cap_bytes, size_t cap_reads)
{
    b.cap_bytes = cap_bytes;
    b.cap_reads = cap_reads;
    if (mode_ == BatchMode::Text) {
        b.seq = (uint8_t *)pinned_get(cap_bytes + 16);
        b.qual = b.has_qual ? (uint8_t *)pinned_get(cap_bytes + 16) : nullptr;
        return;
    }
    b.seq = (uint8_t *)malloc(cap_bytes + 16);
    if (!b.seq) { fprintf(stderr, "kmer_id_b200: out of memory for a %zu-byte batch\n", cap_bytes); exit(1); }
    b.codes = (uint32_t *)pinned_get(4 * batch_code_words(cap_bytes));
    b.boff = (uint32_t *)pinned_get(4 * (cap_reads + 1));
    b.flagbits = (uint32_t *)pinned_get(4 * (cap_reads / 32 + 2));
    b.cap_inv = batch_inv_cap(cap_bytes);
    b.inv = (uint32_t *)pinned_get(4 * b.cap_inv);
    b.taxon = (int32_t *)pinned_get(4 * cap_reads);
}

void ReadBatchReader::free_batch(ReadBatch &b)
{
    if (mode_ == BatchMode::Text) {
        pinned_put(b.seq, b.cap_bytes + 16);
        pinned_put(b.qual, b.cap_bytes + 16);
    } else {
        free(b.seq);
        pinned_put(b.codes, 4 * batch_code_words(b.cap_bytes));
        pinned_put(b.boff, 4 * (b.cap_reads + 1));
        pinned_put(b.flagbits, 4 * (b.cap_reads / 32 + 2));
        pinned_put(b.inv, 4 * b.cap_inv);
        pinned_put(b.taxon, 4 * b.cap_reads);
    }
    b.seq = b.qual = nullptr;
    b.codes = b.boff = b.flagbits = b.inv = nullptr;
    b.taxon = nullptr;
}

ReadBatchReader::ReadBatchReader(ReadFormat fmt, const std::string &path, size_t max_reads, size_t max_bytes, int depth,
                                 unsigned gz_threads, BatchMode mode, unsigned pack_flags)
    : fmt_(fmt), max_reads_(max_reads), max_bytes_(max_bytes), gz_threads_(gz_threads), mode_(mode), pack_flags_(pack_flags)
{
    const bool fastq = fmt == ReadFormat::GzFastq || fmt == ReadFormat::PlainFastq;
    parse_threads_ = fmt == ReadFormat::GzFastq && mode == BatchMode::Packed ? parse_threads_default() : 0;
    // the parallel gz FASTQ path fills one batch per super-block of text (parallel_block_bytes()): size the
    // page-locked buffers for that (ordinary 100-300-base records; the dispatcher grows a batch when a
    // super-block needs more), not for max_bytes of bases - page-locking is slow
    size_t cap_bytes = max_bytes_ + kRefLineLimit, cap_reads = max_reads_;
    if (parse_threads_ > 0) {
        // a super-block is closed by the line block that takes it past the target (<= 4 MiB more)
        const size_t target = parallel_block_bytes(max_bytes_) + std::min<size_t>(4u << 20, parallel_block_bytes(max_bytes_));
        cap_bytes = target / 2 + kRefLineLimit + 64;
        cap_reads = std::max<size_t>(1024, target / 64);
    }
    for (int i = 0; i < depth; i++) {
        auto b = std::make_unique<ReadBatch>();
        b->has_qual = fastq;
        alloc_batch(*b, cap_bytes, cap_reads);
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
    for (auto &b : pool_) free_batch(*b);
}

ReadBatch *ReadBatchReader::get_free()
{
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return !free_.empty(); });
    ReadBatch *b = free_.front();
    free_.pop_front();
    b->n = 0;
    b->n_inv = 0;
    b->n_bases = 0;
    if (b->boff) b->boff[0] = 0;
    b->slot = -1;
    b->span.clear();
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

// trim + pack one record onto the end of a batch's dense arrays (include/kmer_id.h); the caller has
// made sure that bases and reads fit
void ReadBatchReader::pack_record(ReadBatch &b, const char *seq, size_t seqlen, const char *qual)
{
    if (b.n_inv + seqlen > b.cap_inv) { // a record that could overflow the position list (N-rich data): grow it
        const size_t cap = std::max(b.cap_inv * 2, b.n_inv + seqlen + 4096);
        uint32_t *bigger = (uint32_t *)pinned_get(4 * cap);
        memcpy(bigger, b.inv, 4 * b.n_inv);
        pinned_put(b.inv, 4 * b.cap_inv);
        b.inv = bigger;
        b.cap_inv = cap;
    }
    const uint64_t one[2] = { 0, (uint64_t)seqlen };
    uint32_t sp[2], nb = 0;
    size_t ni = 0;
    const size_t w0 = b.n_bases >> 4;
    const int rc = kid_pack_reads_dense((const uint8_t *)seq, (const uint8_t *)qual, one, 1, pack_flags_, b.n_bases,
                                        b.codes + w0, batch_code_words(b.cap_bytes) - w0, b.boff + b.n, b.flagbits, b.n,
                                        b.inv + b.n_inv, b.cap_inv - b.n_inv, &ni, sp, &nb);
    if (rc != 0) { fprintf(stderr, "kmer_id_b200: kid_pack_reads_dense failed (%d)\n", rc); exit(1); }
    b.n_bases = nb;
    b.n_inv += ni;
    b.span.push_back(sp[0]);
    b.span.push_back(sp[1]);
}

// one record into a batch that is known to have room (parallel gz FASTQ: sized per super-block)
void ReadBatchReader::emit_into(ReadBatch &b, const char *acc, size_t acclen, const char *seq, size_t seqlen, const char *qual)
{
    if (b.off.back() + seqlen > b.cap_bytes || b.n >= b.cap_reads) {
        fprintf(stderr, "kmer_id_b200: internal error: a super-block outgrew its batch\n");
        exit(1);
    }
    const uint64_t o = b.off.back();
    memcpy(b.seq + o, seq, seqlen);
    pack_record(b, seq, seqlen, qual);
    b.off.push_back(o + seqlen);
    b.names.insert(b.names.end(), acc, acc + acclen);
    b.name_off.push_back((uint32_t)b.names.size());
    b.n++;
}

// one record as the reference hands it to process_qual (qual != NULL) or process_read (FASTA)
void ReadBatchReader::emit(const char *acc, size_t acclen, const char *seq, size_t seqlen, const char *qual)
{
    ReadBatch *b = cur_;
    const bool packed = mode_ == BatchMode::Packed;
    if (b->off.back() + seqlen > b->cap_bytes || b->n >= b->cap_reads) {
        if (b->n) {
            publish(b);
            b = cur_ = get_free();
        }
        if (seqlen > b->cap_bytes) { // a record longer than a whole batch (a contig): this buffer grows for it
            free_batch(*b);
            alloc_batch(*b, seqlen + kRefLineLimit, max_reads_);
            b->boff[0] = 0;
        }
    }
    const uint64_t o = b->off.back();
    memcpy(b->seq + o, seq, seqlen);
    if (packed) pack_record(*b, seq, seqlen, qual);
    else if (qual) memcpy(b->qual + o, qual, seqlen);
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
    if (fmt_ == ReadFormat::GzFastq && parse_threads_ > 0) {
        run_gz_fastq_parallel(path); // hands out batches in stream order itself, the last one flagged
        return;
    }
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

// ---- gz FASTQ on several threads -----------------------------------------------------------------
// The 4-line state of process_fqgz (:768,:788-802) at any point of the stream is the number of
// NON-EMPTY lines before it, mod 4.  The reader thread only counts those (SSE2, ~10 GB/s), cuts the
// stream into super-blocks of whole line blocks and hands each, with its state and the next few
// lines as look-ahead, to a worker; the worker owns the records whose header line lies in its
// super-block, frames them, trims and packs them (kid_pack_reads) into ONE batch.  Batches are
// handed out in stream order.
namespace {

// non-empty lines of a text that starts at a line start and ends in '\n' (a line that is just "\r" is
// empty too, :786-788)
size_t count_nonempty_lines(const char *p, size_t n)
{
    size_t lines = 0;
    uint64_t prev_nl = 1, prev2_nl = 0, prev_cr = 0; // bytes -1, -2 of the block
    size_t i = 0;
#if defined(__SSE2__)
    const __m128i vnl = _mm_set1_epi8('\n'), vcr = _mm_set1_epi8('\r');
    for (; i + 64 <= n; i += 64) {
        uint64_t nl = 0, cr = 0;
        for (int k = 0; k < 4; k++) {
            const __m128i x = _mm_loadu_si128(reinterpret_cast<const __m128i *>(p + i + 16 * k));
            nl |= (uint64_t)(uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(x, vnl)) << (16 * k);
            cr |= (uint64_t)(uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(x, vcr)) << (16 * k);
        }
        const uint64_t p1 = (nl << 1) | prev_nl, c1 = (cr << 1) | prev_cr;
        const uint64_t p2 = (nl << 2) | (prev_nl << 1) | prev2_nl;
        const uint64_t empty = nl & (p1 | (c1 & p2));
        lines += (size_t)__builtin_popcountll(nl) - (size_t)__builtin_popcountll(empty);
        prev_nl = nl >> 63;
        prev2_nl = (nl >> 62) & 1;
        prev_cr = cr >> 63;
    }
#endif
    for (; i < n; i++) {
        const bool nl = p[i] == '\n';
        if (nl && !(prev_nl || (prev_cr && prev2_nl))) lines++;
        prev2_nl = prev_nl;
        prev_nl = nl;
        prev_cr = p[i] == '\r';
    }
    return lines;
}

struct CountedBlock {
    LineBlock text;
    size_t lines = 0;
};

struct SuperBlock {
    uint64_t index = 0;
    std::vector<LineBlock> blocks, lookahead;
    size_t bytes = 0, lines = 0;
    int state0 = 0; // mod4 at its first line
    bool last = false;
    ReadBatch *batch = nullptr;
};

LineBlock share(const LineBlock &b) // same text, the buffer stays alive through `keep`
{
    LineBlock c;
    c.head = b.head;
    c.body = b.body;
    c.body_len = b.body_len;
    c.keep = b.keep;
    return c;
}

} // namespace

void ReadBatchReader::publish_ordered(uint64_t index, ReadBatch *b)
{
    {
        std::lock_guard<std::mutex> lk(mu_);
        done_[index] = b;
        for (auto it = done_.find(next_pub_); it != done_.end(); it = done_.find(next_pub_)) {
            if (finished_) free_.push_back(it->second); else ready_.push_back(it->second);
            done_.erase(it);
            next_pub_++;
        }
    }
    cv_.notify_all();
}

void ReadBatchReader::run_gz_fastq_parallel(const std::string &path)
{
    const size_t target = parallel_block_bytes(max_bytes_);
    GzLineBlocks src(path, std::min<size_t>(4u << 20, target), gz_threads_);

    std::mutex qmu;
    std::condition_variable qcv;
    std::deque<std::unique_ptr<SuperBlock>> queue;
    bool no_more = false;

    auto work = [&] {
        for (;;) {
            std::unique_ptr<SuperBlock> sb;
            {
                std::unique_lock<std::mutex> lk(qmu);
                qcv.wait(lk, [&] { return !queue.empty() || no_more; });
                if (queue.empty()) return;
                sb = std::move(queue.front());
                queue.pop_front();
            }
            qcv.notify_all();
            ReadBatch *b = sb->batch;
            int mod4 = sb->state0;
            bool have_header = mod4 == 0; // a record whose header lies before this super-block is not ours
            const char *acc = nullptr, *seq = nullptr;
            size_t acclen = 0, seqlen = 0;
            bool stop = false;
            auto lines = [&](const char *p, const char *end, bool look) {
                while (p < end && !stop) {
                    const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
                    size_t len = (size_t)(eol - p);
                    if (len > 0 && p[len - 1] == '\r') len--; // :786-787
                    if (len > 0) {                            // :788 - empty lines do not advance mod4
                        if (mod4 == 0) {
                            if (look) { stop = true; break; } // the next super-block's record
                            acc = p; acclen = len; have_header = true;
                        } else if (mod4 == 1) {
                            seq = p; seqlen = len;
                        } else if (mod4 == 3 && have_header) {
                            if (len < seqlen) // qual.at(stop) throws std::out_of_range (:729): the reference aborts
                                throw std::out_of_range("basic_string::at: quality line shorter than its read");
                            emit_into(*b, acc, acclen, seq, seqlen, p);
                        }
                        mod4 = (mod4 + 1) % 4; // :802
                    }
                    p = eol + 1;
                }
            };
            for (const LineBlock &t : sb->blocks) {
                lines(t.head.data(), t.head.data() + t.head.size(), false);
                lines(t.body, t.body + t.body_len, false);
            }
            if (mod4 != 0 && have_header) // the record that straddles into the next super-block
                for (const LineBlock &t : sb->lookahead) {
                    lines(t.head.data(), t.head.data() + t.head.size(), true);
                    lines(t.body, t.body + t.body_len, true);
                    if (stop || mod4 == 0) break;
                }
            b->last = sb->last;
            const uint64_t index = sb->index;
            sb.reset(); // lets go of the text buffers
            publish_ordered(index, b);
        }
    };
    std::vector<std::thread> workers;
    for (unsigned i = 0; i < parse_threads_; i++) workers.emplace_back(work);

    std::deque<CountedBlock> pend;
    bool eof = false;
    auto pull = [&](size_t want) { // make pend hold `want` blocks if the stream has them
        while (!eof && pend.size() < want) {
            CountedBlock cb;
            if (!src.next(cb.text)) { eof = true; break; }
            cb.lines = count_nonempty_lines(cb.text.head.data(), cb.text.head.size()) +
                       count_nonempty_lines(cb.text.body, cb.text.body_len);
            pend.push_back(std::move(cb));
        }
    };
    int state = 0; // :768
    uint64_t index = 0;
    for (;;) {
        auto sb = std::make_unique<SuperBlock>();
        pull(1);
        while (!pend.empty() && (sb->blocks.empty() || sb->bytes < target)) {
            sb->bytes += pend.front().text.size();
            sb->lines += pend.front().lines;
            sb->blocks.push_back(std::move(pend.front().text));
            pend.pop_front();
            pull(1);
        }
        size_t la = 0;
        for (size_t k = 0; la < 3; k++) { // the straddling record ends within 3 more non-empty lines
            pull(k + 1);
            if (pend.size() <= k) break;
            la += pend[k].lines;
            sb->lookahead.push_back(share(pend[k].text));
        }
        sb->index = index++;
        sb->state0 = state;
        state = (int)((state + sb->lines) % 4);
        sb->last = eof && pend.empty();
        const bool last = sb->last;
        // the batch is granted here, in stream order, so that an earlier super-block never waits for a
        // buffer that later ones hold; every record has >= 4 non-empty lines and its bases are less
        // than half of its text
        ReadBatch *b = get_free();
        const size_t need_reads = sb->lines / 4 + 2, need_bytes = sb->bytes / 2 + kRefLineLimit + 64;
        if (need_reads > b->cap_reads || need_bytes > b->cap_bytes) {
            // rare (an unusually large line block, or very short records): grow in coarse steps so that
            // the page-locked pool keeps handing out buffers of the same few sizes
            auto up = [](size_t v, size_t q) { return (v + v / 4 + q - 1) / q * q; };
            free_batch(*b);
            alloc_batch(*b, need_bytes > b->cap_bytes ? up(need_bytes, (size_t)1 << 20) : b->cap_bytes,
                        need_reads > b->cap_reads ? up(need_reads, (size_t)1 << 16) : b->cap_reads);
        }
        sb->batch = b;
        {
            std::unique_lock<std::mutex> lk(qmu);
            qcv.wait(lk, [&] { return queue.size() < parse_threads_ + 1; });
            queue.push_back(std::move(sb));
        }
        qcv.notify_all();
        if (last) break;
    }
    {
        std::lock_guard<std::mutex> lk(qmu);
        no_more = true;
    }
    qcv.notify_all();
    for (std::thread &w : workers) w.join();
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
