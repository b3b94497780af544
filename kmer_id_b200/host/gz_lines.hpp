// gz_lines.hpp - gzip stream -> '\n'-terminated lines, with the exact edge behaviour of the 16 KiB
// read loop shared by process_kmergz (newkmer_10nx.cpp:675-707) and process_fqgz (:770-810):
//   * a line is everything up to '\n' (a trailing '\r' is the caller's business);
//   * bytes after the last '\n' of the stream are never delivered (:812-813);
//   * a line of >= 16384 bytes is fatal ("Buffer to small for input line lengths", exit 255);
//   * a gz error, or a file that cannot be opened, is fatal (exit 255).
// Unlike the reference this hands out whole blocks of lines, inflated by several threads when the
// file is an ordinary gzip file of some size (pgz.hpp), else by zlib 4 MiB at a time.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <zlib.h>

#include "pgz.hpp"

namespace kidhost {

constexpr size_t kRefLineLimit = 0x4000; // BUFLEN, newkmer_10nx.cpp:85

[[noreturn]] inline void ref_error(const char *msg) // error(), newkmer_10nx.cpp:87-91
{
    fprintf(stderr, "%s\n", msg ? msg : "");
    exit(255);
}

// A run of complete lines without copying the inflated bytes: `head` holds the line that straddles
// the previous run and this one (empty if none), `body` the lines after it inside a buffer that
// `keep` owns.  Every line still ends in '\n'.
struct LineBlock {
    std::vector<char> head;
    const char *body = nullptr;
    size_t body_len = 0;
    std::shared_ptr<void> keep;
    size_t size() const { return head.size() + body_len; }
};

class GzLineBlocks {
public:
    // gz_threads: workers of the parallel inflater (pgz.hpp); 0 = default_gz_threads(), 1 = zlib only
    explicit GzLineBlocks(const std::string &path, size_t block_bytes = 4u << 20, unsigned gz_threads = 0)
        : path_(path), block_(block_bytes)
    {
        pgz_ = ParallelGunzip::open(path, gz_threads);
        if (!pgz_) open_zlib(0);
    }
    ~GzLineBlocks()
    {
        if (in_) gzclose(in_);
    }
    bool parallel() const { return pgz_ != nullptr; }

    // Next run of complete lines.  Returns false at end of stream.  The unterminated tail of the
    // stream, if any, is dropped like the reference does.
    bool next(LineBlock &out)
    {
        out.head.clear();
        out.body = nullptr;
        out.body_len = 0;
        out.keep.reset();
        for (;;) {
            if (eof_) return false;
            const char *d;
            size_t n;
            std::shared_ptr<void> keep;
            if (!fetch(d, n, keep)) {
                eof_ = true;
                return false;
            }
            const char *last = (const char *)memrchr(d, '\n', n);
            if (!last) { // no complete line yet
                carry_.insert(carry_.end(), d, d + n);
                if (carry_.size() >= kRefLineLimit) ref_error("Buffer to small for input line lengths");
                continue;
            }
            const char *body = d;
            if (!carry_.empty()) { // finish the straddling line
                const char *first = (const char *)memchr(d, '\n', n);
                out.head.swap(carry_);
                out.head.insert(out.head.end(), d, first + 1);
                if (out.head.size() > kRefLineLimit) ref_error("Buffer to small for input line lengths");
                body = first + 1;
                carry_.clear();
            }
            out.body = body;
            out.body_len = (size_t)(last + 1 - body);
            out.keep = std::move(keep);
            carry_.assign(last + 1, d + n);
            // every complete line must respect the reference's 16 KiB buffer: hop from a line start
            // to the last '\n' within the next 16 KiB - every line in between is shorter than that
            for (size_t i = 0; i < out.body_len;) {
                const size_t w = std::min(kRefLineLimit, out.body_len - i);
                const char *nl = (const char *)memrchr(body + i, '\n', w);
                if (!nl) ref_error("Buffer to small for input line lengths");
                i = (size_t)(nl - body) + 1;
            }
            if (carry_.size() >= kRefLineLimit) ref_error("Buffer to small for input line lengths");
            return true;
        }
    }
    // Same, copied into one contiguous vector.
    bool next(std::vector<char> &out)
    {
        LineBlock b;
        if (!next(b)) { out.clear(); return false; }
        out.reserve(b.size());
        out.assign(b.head.begin(), b.head.end());
        out.insert(out.end(), b.body, b.body + b.body_len);
        return true;
    }

private:
    void open_zlib(uint64_t skip)
    {
        in_ = gzopen(path_.c_str(), "rb");
        if (!in_) ref_error(nullptr); // gzread(NULL) < 0 -> error(gzerror(NULL)) prints an empty line
        gzbuffer(in_, 1u << 20);
        std::vector<char> scratch(std::min<uint64_t>(skip, 1u << 20));
        while (skip) { // bytes the parallel inflater already delivered
            const int got = gzread(in_, scratch.data(), (unsigned)std::min<uint64_t>(skip, scratch.size()));
            if (got <= 0) break; // the error (if any) shows again on the next read
            skip -= (uint64_t)got;
        }
    }
    // next piece of the inflated stream, owned by `keep`; false at its end
    bool fetch(const char *&d, size_t &n, std::shared_ptr<void> &keep)
    {
        if (pgz_) {
            const uint8_t *p;
            const int rc = pgz_->next(p, n, keep);
            if (rc > 0) { d = (const char *)p; return true; }
            if (rc == 0) return false;
            const uint64_t skip = pgz_->delivered(); // something zlib has to judge: let it
            pgz_.reset();
            open_zlib(skip);
        }
        auto buf = std::make_shared<std::vector<char>>(block_);
        const int got = gzread(in_, buf->data(), (unsigned)buf->size());
        if (got < 0) {
            int err = 0;
            ref_error(gzerror(in_, &err));
        }
        if (got == 0) {
            if (gzclose(in_) != Z_OK) { in_ = nullptr; ref_error("failed gzclose"); }
            in_ = nullptr;
            return false;
        }
        d = buf->data();
        n = (size_t)got;
        keep = buf;
        return true;
    }

    std::string path_;
    size_t block_;
    std::unique_ptr<ParallelGunzip> pgz_;
    gzFile in_ = nullptr;
    std::vector<char> carry_;
    bool eof_ = false;
};

} // namespace kidhost
