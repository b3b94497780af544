// gz_lines.hpp - gzip stream -> '\n'-terminated lines, with the exact edge behaviour of the 16 KiB
// read loop shared by process_kmergz (newkmer_10nx.cpp:675-707) and process_fqgz (:770-810):
//   * a line is everything up to '\n' (a trailing '\r' is the caller's business);
//   * bytes after the last '\n' of the stream are never delivered (:812-813);
//   * a line of >= 16384 bytes is fatal ("Buffer to small for input line lengths", exit 255);
//   * a gz error, or a file that cannot be opened, is fatal (exit 255).
// Unlike the reference this inflates 4 MiB at a time and hands out whole blocks of lines.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <zlib.h>

namespace kidhost {

constexpr size_t kRefLineLimit = 0x4000; // BUFLEN, newkmer_10nx.cpp:85

[[noreturn]] inline void ref_error(const char *msg) // error(), newkmer_10nx.cpp:87-91
{
    fprintf(stderr, "%s\n", msg ? msg : "");
    exit(255);
}

class GzLineBlocks {
public:
    explicit GzLineBlocks(const std::string &path, size_t block_bytes = 4u << 20)
        : block_(block_bytes)
    {
        in_ = gzopen(path.c_str(), "rb");
        if (!in_) ref_error(nullptr); // gzread(NULL) < 0 -> error(gzerror(NULL)) prints an empty line
        gzbuffer(in_, 1u << 20);
        buf_.resize(block_ + kRefLineLimit);
    }
    ~GzLineBlocks()
    {
        if (in_) gzclose(in_);
    }
    // Fills `out` with a run of complete lines (each still ending in '\n').  Returns false at
    // end of stream.  The unterminated tail, if any, is dropped like the reference does.
    bool next(std::vector<char> &out)
    {
        out.clear();
        for (;;) {
            if (eof_) return false;
            const int want = (int)(buf_.size() - have_);
            const int got = gzread(in_, buf_.data() + have_, (unsigned)want);
            if (got < 0) {
                int err = 0;
                ref_error(gzerror(in_, &err));
            }
            if (got == 0) {
                eof_ = true;
                if (gzclose(in_) != Z_OK) { in_ = nullptr; ref_error("failed gzclose"); }
                in_ = nullptr;
                return false;
            }
            const size_t end = have_ + (size_t)got;
            size_t last = end;
            while (last > 0 && buf_[last - 1] != '\n') last--;
            if (last == 0) { // no complete line yet
                if (end >= kRefLineLimit) ref_error("Buffer to small for input line lengths");
                have_ = end;
                continue;
            }
            // every complete line must respect the reference's 16 KiB buffer
            size_t run = 0;
            for (size_t i = 0; i < last; i++) {
                if (buf_[i] == '\n') run = 0;
                else if (++run >= kRefLineLimit) ref_error("Buffer to small for input line lengths");
            }
            if (end - last >= kRefLineLimit) ref_error("Buffer to small for input line lengths");
            out.assign(buf_.begin(), buf_.begin() + (ptrdiff_t)last);
            memmove(buf_.data(), buf_.data() + last, end - last);
            have_ = end - last;
            return true;
        }
    }

private:
    gzFile in_ = nullptr;
    size_t block_;
    std::vector<char> buf_;
    size_t have_ = 0;
    bool eof_ = false;
};

} // namespace kidhost
