// pgz.hpp - multi-threaded inflate of ordinary (single- or multi-member) gzip files.
//
// SURVEY.md 8f row N1 ("parallel inflate"): the reference reads probes10.txt.gz and every FASTQ
// through zlib's gzread on one thread (newkmer_10nx.cpp:675-707, :770-810); with the table on the
// GPU that inflate is what bounds a whole run.  A gzip member cannot be split up front - a deflate
// block may start at any bit and refers to the 32 KiB before it - so this works speculatively:
//
//   * the compressed file is cut into fixed-size pieces; a worker looks, from the first bit of its
//     piece, for something that parses as a block header (dynamic-Huffman header with complete
//     codes, or a gzip member header followed by a block header) and inflates from there to the
//     first block boundary at or after the end of its piece;
//   * it does not know the 32 KiB of history yet, so it inflates to 16-bit symbols: a byte, or a
//     marker "history[i]" (copied around by later matches like any other symbol);
//   * pieces are then chained in file order: a piece is accepted only if it started exactly where
//     its predecessor stopped (so, by induction from the first bit of the file, every accepted
//     start is a real block boundary); its markers are replaced from the predecessor's last 32 KiB
//     and the CRC-32/ISIZE of each finished member are checked.  A piece that started anywhere else
//     (or nowhere) is inflated again from the predecessor's end with known history.
//
// Anything unusual - no gzip magic, trailing garbage, a bad check value, a code this decoder is
// stricter about than zlib, absurd expansion - makes next() return -1; the caller then re-reads the
// file with zlib and skips the bytes already delivered, so error behaviour stays zlib's.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>

namespace kidhost {

class ParallelGunzip {
public:
    // nullptr when the parallel path does not apply (missing/small file, no gzip header, threads < 2).
    // Zero / kDefault arguments take KID_GZ_THREADS, KID_GZ_PIECE_BYTES (1 MiB) and
    // KID_GZ_MIN_BYTES (4 MiB: below that zlib is as fast) from the environment.
    static constexpr size_t kDefault = (size_t)-1;
    static std::unique_ptr<ParallelGunzip> open(const std::string &path, unsigned threads = 0,
                                                size_t piece_bytes = 0, size_t min_file_bytes = kDefault);
    ~ParallelGunzip();
    ParallelGunzip(const ParallelGunzip &) = delete;
    ParallelGunzip &operator=(const ParallelGunzip &) = delete;

    // Next run of inflated bytes in stream order.
    // 1 = data, 0 = clean end of stream, -1 = give up (see above; delivered() says how much to skip).
    // The first form lends the bytes until the next call; the second hands the buffer over (`keep`
    // owns it), so that other threads can work on several runs while this one fetches the next.
    int next(const uint8_t *&data, size_t &len);
    int next(const uint8_t *&data, size_t &len, std::shared_ptr<void> &keep);
    uint64_t delivered() const;
    // pieces accepted as first inflated / inflated again with known history / covered by a predecessor
    void piece_counts(size_t &as_found, size_t &again, size_t &covered) const;

    struct Impl;

private:
    explicit ParallelGunzip(Impl *impl);
    Impl *impl_;
};

// Default worker count: KID_GZ_THREADS if set; else, with h = min(hardware threads, 16): 3h/4 for one
// stream (the probe file, whose lines `h` other threads parse), h / concurrent_files when several
// files are inflated at once (R1 and R2).  Measured on 16 cores: more workers than cores in total
// cost 25-30 % (0.28 s against 0.21 s per 2 M-pair sample).
unsigned default_gz_threads(unsigned concurrent_files = 1);

} // namespace kidhost
