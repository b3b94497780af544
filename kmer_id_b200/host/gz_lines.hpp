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

class GzLineBlocks {
public:
    // gz_threads: workers of the parallel inflater (pgz.hpp); 0 = default_gz_threads(), 1 = zlib only
    explicit GzLineBlocks(const std::string &path, size_t block_bytes = 4u << 20, unsigned gz_threads = 0)
        : path_(path), zbuf_(block_bytes)
    {
        pgz_ = ParallelGunzip::open(path, gz_threads);
        if (!pgz_) open_zlib(0);
    }
    ~GzLineBlocks()
    {
        if (in_) gzclose(in_);
    }
    bool parallel() const { return pgz_ != nullptr; }
    // Fills `out` with a run of complete lines (each still ending in '\n').  Returns false at
    // end of stream.  The unterminated tail, if any, is dropped like the reference does.
    bool next(std::vector<char> &out)
    {
        out.clear();
        for (;;) {
            if (eof_) return false;
            const char *d;
            size_t n;
            if (!fetch(d, n)) {
                eof_ = true;
                return false;
            }
            const char *nl = (const char *)memrchr(d, '\n', n);
            if (!nl) { // no complete line yet
                carry_.insert(carry_.end(), d, d + n);
                if (carry_.size() >= kRefLineLimit) ref_error("Buffer to small for input line lengths");
                continue;
            }
            const size_t upto = (size_t)(nl - d) + 1;
            out.reserve(carry_.size() + upto);
            out.assign(carry_.begin(), carry_.end());
            out.insert(out.end(), d, d + upto);
            carry_.assign(d + upto, d + n);
            // every complete line must respect the reference's 16 KiB buffer: hop from a line start
            // to the last '\n' within the next 16 KiB - every line in between is shorter than that
            for (size_t i = 0; i < out.size();) {
                const size_t w = std::min(kRefLineLimit, out.size() - i);
                const char *last = (const char *)memrchr(out.data() + i, '\n', w);
                if (!last) ref_error("Buffer to small for input line lengths");
                i = (size_t)(last - out.data()) + 1;
            }
            if (carry_.size() >= kRefLineLimit) ref_error("Buffer to small for input line lengths");
            return true;
        }
    }

private:
    void open_zlib(uint64_t skip)
    {
        in_ = gzopen(path_.c_str(), "rb");
        if (!in_) ref_error(nullptr); // gzread(NULL) < 0 -> error(gzerror(NULL)) prints an empty line
        gzbuffer(in_, 1u << 20);
        while (skip) { // bytes the parallel inflater already delivered
            const int got = gzread(in_, zbuf_.data(), (unsigned)std::min<uint64_t>(skip, zbuf_.size()));
            if (got <= 0) break; // the error (if any) shows again on the next read
            skip -= (uint64_t)got;
        }
    }
    // next piece of the inflated stream; false at its end
    bool fetch(const char *&d, size_t &n)
    {
        if (pgz_) {
            const uint8_t *p;
            const int rc = pgz_->next(p, n);
            if (rc > 0) { d = (const char *)p; return true; }
            if (rc == 0) return false;
            const uint64_t skip = pgz_->delivered(); // something zlib has to judge: let it
            pgz_.reset();
            open_zlib(skip);
        }
        const int got = gzread(in_, zbuf_.data(), (unsigned)zbuf_.size());
        if (got < 0) {
            int err = 0;
            ref_error(gzerror(in_, &err));
        }
        if (got == 0) {
            if (gzclose(in_) != Z_OK) { in_ = nullptr; ref_error("failed gzclose"); }
            in_ = nullptr;
            return false;
        }
        d = zbuf_.data();
        n = (size_t)got;
        return true;
    }

    std::string path_;
    std::unique_ptr<ParallelGunzip> pgz_;
    gzFile in_ = nullptr;
    std::vector<char> zbuf_, carry_;
    bool eof_ = false;
};

} // namespace kidhost
