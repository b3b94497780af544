// pgz.cpp - see pgz.hpp.  Own deflate decoder (RFC 1951) because zlib can neither start at an
// arbitrary bit without the 32 KiB of history nor carry "unknown history" markers through matches.
// Wherever zlib's inflate would refuse a stream this decoder refuses it too (or earlier): every
// refusal ends in the caller's zlib fallback, never in different data.
#include "pgz.hpp"

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fcntl.h>
#include <mutex>
#include <sys/mman.h>
#include <sys/resource.h>
#include <sys/syscall.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>
#include <zlib.h>

namespace kidhost {

unsigned default_gz_threads(unsigned concurrent_files)
{
    if (const char *e = getenv("KID_GZ_THREADS")) return (unsigned)std::max(0, atoi(e));
    const unsigned h = std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
    if (h < 2) return 1; // zlib
    return std::max(2u, concurrent_files > 1 ? h / concurrent_files : 3 * h / 4);
}

namespace {

constexpr int kWin = 32768;    // deflate history
constexpr int kLitRoot = 10;   // first-level bits of the literal/length table
constexpr int kDistRoot = 8;
constexpr int kClRoot = 7;
constexpr uint32_t kMaxExpand = 64; // refuse pieces that inflate to more than this many times their size

// ---- decode-table entries --------------------------------------------------------------------
// bits 0-4 code bits to drop | 5-7 kind | 8-12 extra bits (or second-level bits) | 16-31 value
enum Kind : uint32_t { K_LIT = 0, K_LEN = 1, K_EOB = 2, K_SUB = 3, K_BAD = 4 };
constexpr uint32_t entry(Kind k, uint32_t extra, uint32_t val) { return ((uint32_t)k << 5) | (extra << 8) | (val << 16); }
inline uint32_t e_bits(uint32_t e) { return e & 31u; }
inline uint32_t e_kind(uint32_t e) { return (e >> 5) & 7u; }
inline uint32_t e_extra(uint32_t e) { return (e >> 8) & 31u; }
inline uint32_t e_val(uint32_t e) { return e >> 16; }

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct Payloads {
    uint32_t lit[288], dist[32], cl[19];
    Payloads()
    {
        for (int s = 0; s < 256; s++) lit[s] = entry(K_LIT, 0, (uint32_t)s);
        lit[256] = entry(K_EOB, 0, 0);
        for (int s = 257; s < 286; s++) lit[s] = entry(K_LEN, kLenExtra[s - 257], kLenBase[s - 257]);
        lit[286] = lit[287] = entry(K_BAD, 0, 0);
        for (int s = 0; s < 30; s++) dist[s] = entry(K_LEN, kDistExtra[s], kDistBase[s]);
        dist[30] = dist[31] = entry(K_BAD, 0, 0);
        for (int s = 0; s < 19; s++) cl[s] = entry(K_LIT, 0, (uint32_t)s);
    }
};
const Payloads kPay;

enum class CodeUse { LitLen, Dist, CodeLen };

inline uint32_t reverse_bits(uint32_t c, int n)
{
    uint32_t r = 0;
    for (int i = 0; i < n; i++) { r = (r << 1) | (c & 1u); c >>= 1; }
    return r;
}

// Canonical Huffman code -> two-level table.  false = zlib's inflate_table would reject the code
// (over-subscribed, or incomplete where inflate does not allow it) or we are stricter.
bool build_table(const uint8_t *lens, int n, int root, const uint32_t *payload, CodeUse use, std::vector<uint32_t> &tab)
{
    int count[16] = {0};
    for (int s = 0; s < n; s++) count[lens[s]]++;
    const int used = n - count[0];
    tab.assign((size_t)1 << root, entry(K_BAD, 0, 0) | 1u);
    if (used == 0) return use == CodeUse::Dist; // no distance codes: legal, any match is an error
    int left = 1, maxlen = 0;
    for (int l = 1; l <= 15; l++) {
        left = (left << 1) - count[l];
        if (left < 0) return false;
        if (count[l]) maxlen = l;
    }
    if (left > 0 && !(use == CodeUse::Dist && used == 1 && maxlen == 1)) return false;
    uint32_t next[16];
    uint32_t code = 0;
    count[0] = 0;
    for (int l = 1; l <= 15; l++) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
    if (maxlen > root) { // size of each second-level table = longest code under its first-level prefix
        uint8_t sub[1 << kLitRoot];
        memset(sub, 0, (size_t)1 << root);
        uint32_t nx[16];
        memcpy(nx, next, sizeof nx);
        for (int s = 0; s < n; s++) {
            const int l = lens[s];
            if (!l) continue;
            const uint32_t r = reverse_bits(nx[l]++, l);
            if (l > root) { uint8_t &m = sub[r & ((1u << root) - 1)]; m = std::max<uint8_t>(m, (uint8_t)(l - root)); }
        }
        for (uint32_t p = 0; p < (1u << root); p++)
            if (sub[p]) {
                const size_t off = tab.size();
                if (off + ((size_t)1 << sub[p]) > 65535) return false;
                tab.resize(off + ((size_t)1 << sub[p]), entry(K_BAD, 0, 0) | 1u);
                tab[p] = entry(K_SUB, sub[p], (uint32_t)off) | (uint32_t)root;
            }
    }
    for (int s = 0; s < n; s++) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t r = reverse_bits(next[l]++, l);
        if (l <= root) {
            for (uint32_t i = r; i < (1u << root); i += 1u << l) tab[i] = payload[s] | (uint32_t)l;
        } else {
            const uint32_t head = tab[r & ((1u << root) - 1)];
            const uint32_t off = e_val(head), sb = e_extra(head);
            for (uint32_t i = r >> root; i < (1u << sb); i += 1u << (l - root)) tab[off + i] = payload[s] | (uint32_t)(l - root);
        }
    }
    return true;
}

// ---- bit reader (LSB first) ------------------------------------------------------------------
struct Bits {
    const uint8_t *base = nullptr, *p = nullptr, *end = nullptr;
    uint64_t buf = 0;
    int cnt = 0;
    void seek(const uint8_t *b, const uint8_t *e, uint64_t bitpos)
    {
        base = b; end = e; p = b + (bitpos >> 3); buf = 0; cnt = 0;
        refill();
        drop((int)(bitpos & 7));
    }
    inline void refill()
    {
        if (p + 8 <= end) {
            uint64_t v;
            memcpy(&v, p, 8);
            buf |= v << cnt;
            p += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56) { // past the end: zeros, noticed through pos() > size
                const uint64_t b = p < end ? *p : 0;
                buf |= b << cnt;
                p++;
                cnt += 8;
            }
        }
    }
    inline void drop(int n) { buf >>= n; cnt -= n; }
    inline uint32_t take(int n)
    {
        const uint32_t v = (uint32_t)(buf & ((1ull << n) - 1));
        drop(n);
        return v;
    }
    uint64_t pos() const { return (uint64_t)(p - base) * 8 - (uint64_t)cnt; }
};

struct Codes {
    std::vector<uint32_t> lit, dist;
};

const Codes &fixed_codes()
{
    static const Codes &c = *[] { // leaked on purpose, see g_pool
        Codes &f = *new Codes;
        uint8_t l[288];
        for (int s = 0; s < 288; s++) l[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
        build_table(l, 288, kLitRoot, kPay.lit, CodeUse::LitLen, f.lit);
        uint8_t d[32];
        memset(d, 5, sizeof d);
        build_table(d, 32, kDistRoot, kPay.dist, CodeUse::Dist, f.dist);
        return &f;
    }();
    return c;
}

// Dynamic block header (after the 3 block-type bits).  Mirrors inflate's checks.
bool read_dynamic_header(Bits &in, Codes &c, std::vector<uint32_t> &cltab)
{
    in.refill();
    const uint32_t nlen = in.take(5) + 257, ndist = in.take(5) + 1, ncl = in.take(4) + 4;
    if (nlen > 286 || ndist > 30) return false;
    uint8_t cl[19] = {0};
    in.refill();
    for (uint32_t i = 0; i < ncl; i++) {
        if (in.cnt < 3) in.refill();
        cl[kClOrder[i]] = (uint8_t)in.take(3);
    }
    if (!build_table(cl, 19, kClRoot, kPay.cl, CodeUse::CodeLen, cltab)) return false;
    uint8_t lens[286 + 30];
    uint32_t have = 0;
    while (have < nlen + ndist) {
        in.refill();
        const uint32_t e = cltab[in.buf & ((1u << kClRoot) - 1)];
        if (e_kind(e) != K_LIT) return false;
        in.drop((int)e_bits(e));
        const uint32_t sym = e_val(e);
        if (sym < 16) { lens[have++] = (uint8_t)sym; continue; }
        uint32_t rep;
        uint8_t val = 0;
        if (sym == 16) {
            if (have == 0) return false;
            val = lens[have - 1];
            rep = 3 + in.take(2);
        } else if (sym == 17) rep = 3 + in.take(3);
        else rep = 11 + in.take(7);
        if (have + rep > nlen + ndist) return false;
        memset(lens + have, val, rep);
        have += rep;
    }
    if (lens[256] == 0) return false; // no end-of-block code
    return build_table(lens, (int)nlen, kLitRoot, kPay.lit, CodeUse::LitLen, c.lit) &&
           build_table(lens + nlen, (int)ndist, kDistRoot, kPay.dist, CodeUse::Dist, c.dist);
}

// ---- gzip framing ----------------------------------------------------------------------------
// Member header at byte `at`; returns its length, 0 if it is not a header we accept.
size_t gzip_header_len(const uint8_t *d, size_t size, size_t at)
{
    if (at + 10 > size || d[at] != 0x1f || d[at + 1] != 0x8b || d[at + 2] != 8 || (d[at + 3] & 0xe0)) return 0;
    const uint8_t flg = d[at + 3];
    size_t p = at + 10;
    if (flg & 4) { // FEXTRA
        if (p + 2 > size) return 0;
        p += 2 + ((size_t)d[p] | ((size_t)d[p + 1] << 8));
    }
    for (int f = 8; f <= 16; f <<= 1) // FNAME, FCOMMENT: zero-terminated
        if (flg & f) {
            while (p < size && d[p]) p++;
            p++;
        }
    if (flg & 2) p += 2; // FHCRC
    return p < size ? p - at : 0;
}

// ---- one piece of output -----------------------------------------------------------------------
struct MemberEnd {
    size_t out_pos;  // symbols of this piece that belong to the member that ends here
    uint32_t crc, isize;
};

// Piece buffers are megabytes each and live for milliseconds; handing them back to malloc means an
// mmap/munmap pair and a page fault per 4 KiB every time, so they are recycled here instead.
class BlockPool {
public:
    void *get(size_t bytes, size_t &got)
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (size_t i = 0; i < free_.size(); i++)
                if (free_[i].second >= bytes && free_[i].second <= 4 * bytes + (1u << 20)) {
                    void *p = free_[i].first;
                    got = free_[i].second;
                    free_[i] = free_.back();
                    free_.pop_back();
                    return p;
                }
        }
        got = bytes;
        return malloc(bytes);
    }
    void put(void *p, size_t bytes)
    {
        if (!p) return;
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (free_.size() < 96) { free_.emplace_back(p, bytes); return; }
        }
        free(p);
    }
    ~BlockPool()
    {
        for (auto &b : free_) free(b.first);
    }

private:
    std::mutex mu_;
    std::vector<std::pair<void *, size_t>> free_;
};
// Never destroyed: ref_error() may exit() while workers still decode into these buffers.
BlockPool &g_pool = *new BlockPool;

struct SymBuf { // kWin symbols of history, then the output
    uint16_t *p = nullptr;
    size_t cap = 0; // symbols
    ~SymBuf() { release(); }
    bool reserve(size_t n)
    {
        if (n <= cap) return true;
        size_t got;
        uint16_t *q = (uint16_t *)g_pool.get(n * sizeof(uint16_t), got);
        if (!q) return false;
        if (p) memcpy(q, p, cap * sizeof(uint16_t));
        g_pool.put(p, cap * sizeof(uint16_t));
        p = q;
        cap = got / sizeof(uint16_t);
        return true;
    }
    void release() { g_pool.put(p, cap * sizeof(uint16_t)); p = nullptr; cap = 0; }
};

struct Run {
    SymBuf sym;
    size_t n = 0; // output symbols (after the kWin prefix)
    std::vector<MemberEnd> ends;
    uint64_t start_bit = 0, end_bit = 0;
    bool eof = false;
};

// Inflates blocks from in.pos() (a block header) until a block header at or after stop_bit, or the
// clean end of the file.  floor = lowest index of run.sym a match may reach.  false = refuse.
bool inflate_run(const uint8_t *data, size_t size, Bits &in, Run &run, size_t floor, uint64_t stop_bit, size_t max_out)
{
    const uint64_t size_bits = (uint64_t)size * 8;
    Codes dyn;
    std::vector<uint32_t> cltab;
    uint16_t *out = run.sym.p;
    size_t pos = kWin + run.n, cap = run.sym.cap;
    auto room = [&](size_t need) {
        if (pos + need <= cap) return true;
        if (pos - kWin + need > max_out) return false;
        if (!run.sym.reserve(std::max(cap * 2, pos + need))) return false;
        out = run.sym.p;
        cap = run.sym.cap;
        return true;
    };
    for (;;) {
        const uint64_t here = in.pos();
        if (here >= stop_bit) { run.end_bit = here; run.n = pos - kWin; return here <= size_bits; }
        if (here + 3 > size_bits) return false;
        in.refill();
        const uint32_t bfinal = in.take(1), btype = in.take(2);
        if (btype == 3) return false;
        if (btype == 0) {
            size_t at = (size_t)((in.pos() + 7) >> 3);
            if (at + 4 > size) return false;
            const uint32_t len = data[at] | (data[at + 1] << 8), nlen = data[at + 2] | (data[at + 3] << 8);
            if ((len ^ 0xffffu) != nlen) return false;
            at += 4;
            if (at + len > size || !room(len + 8)) return false;
            for (uint32_t i = 0; i < len; i++) out[pos + i] = data[at + i];
            pos += len;
            in.seek(data, data + size, (uint64_t)(at + len) * 8);
        } else {
            const Codes *c = &fixed_codes();
            if (btype == 2) {
                if (!read_dynamic_header(in, dyn, cltab)) return false;
                c = &dyn;
            }
            const uint32_t *lit = c->lit.data(), *dist = c->dist.data();
            Bits b = in; // a local copy: the byte copies below may alias anything but this
            for (;;) {
                if (pos + 320 > cap && !room(320)) return false;
                b.refill();
                if (b.p > b.end + 64) return false; // ran off the end of the file
                uint32_t e = lit[b.buf & ((1u << kLitRoot) - 1)];
                if (e_kind(e) == K_SUB) {
                    b.drop(kLitRoot);
                    e = lit[e_val(e) + (b.buf & ((1u << e_extra(e)) - 1))];
                }
                b.drop((int)e_bits(e));
                const uint32_t k = e_kind(e);
                if (k == K_LIT) {
                    out[pos++] = (uint16_t)e_val(e);
                    // a second literal from the same refill: >= 41 bits are left
                    e = lit[b.buf & ((1u << kLitRoot) - 1)];
                    if (e_kind(e) == K_LIT) {
                        b.drop((int)e_bits(e));
                        out[pos++] = (uint16_t)e_val(e);
                    }
                    continue;
                }
                if (k == K_LEN) {
                    const uint32_t len = e_val(e) + b.take((int)e_extra(e));
                    uint32_t d = dist[b.buf & ((1u << kDistRoot) - 1)];
                    if (e_kind(d) == K_SUB) {
                        b.drop(kDistRoot);
                        d = dist[e_val(d) + (b.buf & ((1u << e_extra(d)) - 1))];
                    }
                    if (e_kind(d) != K_LEN) return false;
                    b.drop((int)e_bits(d));
                    const size_t back = e_val(d) + b.take((int)e_extra(d));
                    if (back > pos - floor) return false; // before the member / before any history
                    uint16_t *dst = out + pos;
                    const uint16_t *src = dst - back;
                    if (back >= 8) { // 8 symbols at a time; most matches end with the first copy
                        memcpy(dst, src, 16);
                        for (uint32_t i = 8; i < len; i += 8) memcpy(dst + i, src + i, 16);
                    } else { // overlapping run: the output is periodic with period `back`
                        const uint32_t head = len < 16 ? len : 16;
                        for (uint32_t i = 0; i < head; i++) dst[i] = src[i];
                        if (len > 16) {
                            const size_t eff = back * ((7 + back) / back); // multiple of the period, 8..14
                            for (uint32_t i = 16; i < len; i += 8) memcpy(dst + i, dst + i - eff, 16);
                        }
                    }
                    pos += len;
                    continue;
                }
                if (k == K_EOB) break;
                return false;
            }
            in = b;
            if (in.pos() > size_bits) return false;
        }
        if (bfinal) {
            size_t at = (size_t)((in.pos() + 7) >> 3);
            if (at + 8 > size) return false;
            MemberEnd me;
            me.out_pos = pos - kWin;
            me.crc = (uint32_t)data[at] | ((uint32_t)data[at + 1] << 8) | ((uint32_t)data[at + 2] << 16) | ((uint32_t)data[at + 3] << 24);
            me.isize = (uint32_t)data[at + 4] | ((uint32_t)data[at + 5] << 8) | ((uint32_t)data[at + 6] << 16) | ((uint32_t)data[at + 7] << 24);
            run.ends.push_back(me);
            at += 8;
            if (at == size) { run.end_bit = size_bits; run.eof = true; run.n = pos - kWin; return true; }
            const size_t hl = gzip_header_len(data, size, at);
            if (!hl) return false; // trailing bytes that are not another member: zlib's business
            floor = pos;
            in.seek(data, data + size, (uint64_t)(at + hl) * 8);
        }
    }
}

inline uint64_t peek64(const uint8_t *d, uint64_t bit)
{
    uint64_t v;
    memcpy(&v, d + (bit >> 3), 8);
    return v >> (bit & 7);
}

// First bit position in [from, to) that passes as a block start: a dynamic-Huffman block with
// complete codes, or (byte aligned) a gzip member header (then start = its first block header).
// cand = the position that matched, to resume the search from if the candidate does not inflate.
bool find_block(const uint8_t *data, size_t size, uint64_t from, uint64_t to, uint64_t &cand, uint64_t &start, bool &at_member)
{
    if (size < 16) return false;
    to = std::min<uint64_t>(to, (uint64_t)(size - 16) * 8);
    Codes tmp;
    std::vector<uint32_t> cltab;
    for (uint64_t b = from; b < to; b++) {
        if ((b & 7) == 0) {
            const size_t at = (size_t)(b >> 3);
            if (data[at] == 0x1f && data[at + 1] == 0x8b) {
                const size_t hl = gzip_header_len(data, size, at);
                if (hl) { cand = b; start = (uint64_t)(at + hl) * 8; at_member = true; return true; }
            }
        }
        const uint64_t v = peek64(data, b);
        if ((v & 6) != 4) continue; // BTYPE = 2 (dynamic codes), last block or not
        if (((v >> 3) & 31) > 29 || ((v >> 8) & 31) > 29) continue;
        const uint32_t ncl = (uint32_t)((v >> 13) & 15) + 4;
        uint64_t w = peek64(data, b + 17);
        uint32_t kraft = 0;
        for (uint32_t i = 0; i < ncl; i++, w >>= 3) {
            const uint32_t l = (uint32_t)(w & 7);
            if (l) kraft += 128u >> l;
        }
        if (kraft != 128) continue;
        Bits in;
        in.seek(data, data + size, b + 3);
        if (!read_dynamic_header(in, tmp, cltab)) continue;
        cand = start = b;
        at_member = false;
        return true;
    }
    return false;
}

enum State { IDLE, DECODING, DECODED, WINDOWED, RESOLVING, READY, FAILED };

struct Segment {
    size_t len;
    uint32_t crc;
    bool ends_member;
    uint32_t want_crc, want_isize;
};

struct Piece {
    State state = IDLE;
    bool found = false;
    Run run;
    std::vector<uint8_t> window; // history this piece's markers refer to (set by the chain)
    size_t window_valid = 0;
    uint8_t *bytes = nullptr;
    size_t n_bytes = 0, cap_bytes = 0;
    std::vector<Segment> segs;
    void drop_bytes() { g_pool.put(bytes, cap_bytes); bytes = nullptr; cap_bytes = 0; }
    ~Piece() { drop_bytes(); }
};

} // namespace

struct ParallelGunzip::Impl {
    const uint8_t *data = nullptr;
    size_t size = 0;
    int fd = -1;
    size_t piece_bytes = 0;
    size_t n_pieces = 0;
    uint64_t first_block_bit = 0;
    std::vector<Piece> pieces;

    std::mutex mu;
    std::condition_variable cv;
    std::vector<std::thread> pool;
    std::deque<size_t> resolve_q;
    size_t next_decode = 0, consumed = 0, max_ahead = 0;
    bool stop = false;

    // chain state (owned by whoever holds chain_busy)
    bool chain_busy = false;
    size_t chain_next = 0;
    uint64_t chain_end_bit = 0;
    bool chain_eof = false;
    size_t n_spec = 0, n_redo = 0, n_skip = 0; // pieces accepted as inflated / inflated again / covered
    std::vector<uint8_t> chain_window = std::vector<uint8_t>(kWin, 0);
    size_t chain_valid = 0;

    // consumer state
    bool failed = false;
    uint64_t delivered = 0;
    uint32_t run_crc = 0;
    uint64_t run_len = 0;
    bool in_member_data = false;
    std::shared_ptr<void> lent; // buffer behind the pointer the lending next() returned

    uint64_t piece_end_bit(size_t k) const { return (uint64_t)std::min(size, (k + 1) * piece_bytes) * 8; }
    size_t max_out() const { return (size_t)kMaxExpand * piece_bytes + (1u << 20); }

    bool init_unknown_history(Run &r)
    {
        r.n = 0;
        r.ends.clear();
        r.eof = false;
        if (!r.sym.reserve(kWin + 6 * piece_bytes)) return false; // out of memory: the piece is "not found"
        for (int i = 0; i < kWin; i++) r.sym.p[i] = (uint16_t)(256 + i);
        return true;
    }

    // speculative decode of piece k (k = 0 starts at the known first block)
    void decode_piece(size_t k)
    {
        Piece &pc = pieces[k];
        const uint64_t stop_bit = piece_end_bit(k);
        if (k == 0) {
            if (!init_unknown_history(pc.run)) { pc.found = false; return; }
            Bits in;
            in.seek(data, data + size, first_block_bit);
            pc.run.start_bit = first_block_bit;
            pc.found = inflate_run(data, size, in, pc.run, kWin, stop_bit, max_out());
            return;
        }
        uint64_t from = (uint64_t)k * piece_bytes * 8;
        for (int tries = 0; tries < 64; tries++) {
            uint64_t cand, start;
            bool at_member;
            if (!find_block(data, size, from, stop_bit, cand, start, at_member)) break;
            if (!init_unknown_history(pc.run)) break;
            Bits in;
            in.seek(data, data + size, start);
            pc.run.start_bit = start;
            if (inflate_run(data, size, in, pc.run, at_member ? kWin : 0, stop_bit, max_out())) {
                pc.found = true;
                return;
            }
            from = cand + 1; // not a block after all
        }
        pc.found = false;
        pc.run.sym.release();
    }

    // piece k did not start where its predecessor ended: inflate again with the real history
    bool redo_with_history(size_t k)
    {
        Piece &pc = pieces[k];
        Run &r = pc.run;
        r.n = 0;
        r.ends.clear();
        r.eof = false;
        if (!r.sym.reserve(kWin + 6 * piece_bytes)) return false;
        for (int i = 0; i < kWin; i++) r.sym.p[i] = chain_window[i];
        r.start_bit = chain_end_bit;
        Bits in;
        in.seek(data, data + size, chain_end_bit);
        return inflate_run(data, size, in, r, kWin - chain_valid, piece_end_bit(k), max_out());
    }

    // returns false if piece k cannot be produced by this decoder
    bool chain_piece(size_t k)
    {
        Piece &pc = pieces[k];
        Run &r = pc.run;
        if (chain_eof || (k > 0 && chain_end_bit >= piece_end_bit(k))) { // already covered by a predecessor
            r.sym.release();
            r.n = 0;
            r.ends.clear();
            pc.window_valid = 0;
            n_skip++;
            return true;
        }
        if (!(pc.found && r.start_bit == chain_end_bit)) {
            if (k == 0 || !redo_with_history(k)) return false;
            n_redo++;
        } else n_spec++;
        pc.window = chain_window;
        pc.window_valid = chain_valid;
        // history for the successor: last kWin bytes of (window ++ output), markers resolved
        const size_t keep = std::min<size_t>(r.n, kWin);
        std::vector<uint8_t> nw(kWin);
        if (keep < (size_t)kWin) memcpy(nw.data(), chain_window.data() + keep, kWin - keep);
        const uint16_t *tail = r.sym.p + kWin + r.n - keep;
        for (size_t i = 0; i < keep; i++) {
            const uint16_t s = tail[i];
            if (s < 256) nw[kWin - keep + i] = (uint8_t)s;
            else {
                if ((size_t)(s - 256) < kWin - chain_valid) return false; // refers to before its member
                nw[kWin - keep + i] = chain_window[s - 256];
            }
        }
        const size_t since_member = r.ends.empty() ? chain_valid + r.n : r.n - r.ends.back().out_pos;
        chain_window.swap(nw);
        chain_valid = std::min<size_t>(since_member, kWin);
        chain_end_bit = r.end_bit;
        chain_eof = r.eof;
        return true;
    }

    void advance_chain()
    {
        std::unique_lock<std::mutex> lk(mu);
        while (!chain_busy && chain_next < n_pieces && pieces[chain_next].state == DECODED) {
            chain_busy = true;
            const size_t k = chain_next;
            lk.unlock();
            const bool ok = chain_piece(k);
            lk.lock();
            pieces[k].state = ok ? WINDOWED : FAILED;
            if (ok) resolve_q.push_back(k);
            else chain_next = next_decode = n_pieces; // nothing after a failure is ever needed
            if (ok) chain_next++;
            chain_busy = false;
            cv.notify_all();
        }
    }

    bool resolve_piece(size_t k)
    {
        Piece &pc = pieces[k];
        Run &r = pc.run;
        pc.n_bytes = r.n;
        pc.bytes = (uint8_t *)g_pool.get(std::max<size_t>(r.n, 64), pc.cap_bytes);
        if (!pc.bytes) return false;
        if (r.n) {
            uint8_t lut[256 + kWin];
            for (int i = 0; i < 256; i++) lut[i] = (uint8_t)i;
            memcpy(lut + 256, pc.window.data(), kWin);
            const uint16_t *s = r.sym.p + kWin;
            if (pc.window_valid < (size_t)kWin) {
                const uint32_t lowest = 256 + (uint32_t)(kWin - pc.window_valid);
                for (size_t i = 0; i < r.n; i++)
                    if (s[i] >= 256 && s[i] < lowest) return false;
            }
            // bytes and their CRC-32 in one sweep, 64 KiB at a time, so that the CRC reads the cache
            constexpr size_t kTile = 64u << 10;
            size_t from = 0;
            auto sweep = [&](size_t to, bool ends_member, uint32_t want_crc, uint32_t want_isize) {
                uint32_t crc = 0;
                for (size_t a = from; a < to; a += kTile) {
                    const size_t b = std::min(to, a + kTile);
                    for (size_t i = a; i < b; i++) pc.bytes[i] = lut[s[i]];
                    crc = (uint32_t)crc32_z(crc, pc.bytes + a, b - a);
                }
                pc.segs.push_back(Segment{to - from, crc, ends_member, want_crc, want_isize});
                from = to;
            };
            for (const MemberEnd &me : r.ends) sweep(me.out_pos, true, me.crc, me.isize);
            if (from < r.n) sweep(r.n, false, 0, 0);
        } else {
            for (const MemberEnd &me : r.ends) pc.segs.push_back(Segment{0, 0, true, me.crc, me.isize});
        }
        r.sym.release();
        std::vector<uint8_t>().swap(pc.window);
        return true;
    }

    void worker()
    {
#ifdef __linux__
        // The consumer of the inflated text (one parser thread per file) is the serial stage of the
        // pipeline; when the machine is oversubscribed it, not these workers, should get the core.
        setpriority(PRIO_PROCESS, (id_t)syscall(SYS_gettid), 5);
#endif
        for (;;) {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return stop || !resolve_q.empty() || (next_decode < n_pieces && next_decode < consumed + max_ahead); });
            if (stop) return;
            if (!resolve_q.empty()) {
                const size_t k = resolve_q.front();
                resolve_q.pop_front();
                pieces[k].state = RESOLVING;
                lk.unlock();
                const bool ok = resolve_piece(k);
                lk.lock();
                pieces[k].state = ok ? READY : FAILED;
                cv.notify_all();
                continue;
            }
            const size_t k = next_decode++;
            pieces[k].state = DECODING;
            lk.unlock();
            decode_piece(k);
            lk.lock();
            pieces[k].state = DECODED;
            lk.unlock();
            advance_chain();
        }
    }

    ~Impl()
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (auto &t : pool) t.join();
        if (data) munmap((void *)data, size);
        if (fd >= 0) close(fd);
    }
};

ParallelGunzip::ParallelGunzip(Impl *impl) : impl_(impl) {}
ParallelGunzip::~ParallelGunzip() { delete impl_; }
uint64_t ParallelGunzip::delivered() const { return impl_->delivered; }
void ParallelGunzip::piece_counts(size_t &as_found, size_t &again, size_t &covered) const
{
    std::lock_guard<std::mutex> lk(impl_->mu);
    as_found = impl_->n_spec;
    again = impl_->n_redo;
    covered = impl_->n_skip;
}

std::unique_ptr<ParallelGunzip> ParallelGunzip::open(const std::string &path, unsigned threads, size_t piece_bytes, size_t min_file_bytes)
{
    auto env_size = [](const char *name, size_t def) {
        const char *e = getenv(name);
        return e && *e ? (size_t)strtoull(e, nullptr, 10) : def;
    };
    if (threads == 0) threads = default_gz_threads();
    if (piece_bytes == 0) piece_bytes = env_size("KID_GZ_PIECE_BYTES", (size_t)1 << 20);
    if (min_file_bytes == kDefault) min_file_bytes = env_size("KID_GZ_MIN_BYTES", (size_t)4 << 20);
    if (threads < 2 || piece_bytes < 1024) return nullptr;
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return nullptr;
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || (size_t)st.st_size < std::max<size_t>(min_file_bytes, 64)) {
        close(fd);
        return nullptr;
    }
    void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) {
        close(fd);
        return nullptr;
    }
    madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
    auto *im = new Impl;
    im->data = (const uint8_t *)m;
    im->size = (size_t)st.st_size;
    im->fd = fd;
    const size_t hl = gzip_header_len(im->data, im->size, 0);
    if (!hl) {
        delete im;
        return nullptr;
    }
    im->first_block_bit = (uint64_t)hl * 8;
    im->chain_end_bit = im->first_block_bit;
    im->piece_bytes = piece_bytes;
    im->n_pieces = (im->size + piece_bytes - 1) / piece_bytes;
    im->pieces = std::vector<Piece>(im->n_pieces);
    im->max_ahead = 2 * (size_t)threads + 2;
    for (unsigned t = 0; t < threads; t++) im->pool.emplace_back([im] { im->worker(); });
    return std::unique_ptr<ParallelGunzip>(new ParallelGunzip(im));
}

int ParallelGunzip::next(const uint8_t *&out, size_t &len)
{
    return next(out, len, impl_->lent); // the previous run goes back to the pool here
}

int ParallelGunzip::next(const uint8_t *&out, size_t &len, std::shared_ptr<void> &keep)
{
    Impl &s = *impl_;
    keep.reset();
    if (s.failed) return -1;
    for (;;) {
        std::unique_lock<std::mutex> lk(s.mu);
        if (s.consumed == s.n_pieces) {
            if (!s.chain_eof || s.in_member_data) { s.failed = true; return -1; }
            return 0;
        }
        const size_t k = s.consumed;
        s.cv.wait(lk, [&] { return s.pieces[k].state == READY || s.pieces[k].state == FAILED; });
        if (s.pieces[k].state == FAILED) { s.failed = true; return -1; }
        lk.unlock();
        Piece &pc = s.pieces[k];
        for (const Segment &sg : pc.segs) { // check values of the members that end in this piece
            s.run_crc = (uint32_t)crc32_combine(s.run_crc, sg.crc, (z_off_t)sg.len);
            s.run_len += sg.len;
            s.in_member_data = true;
            if (sg.ends_member) {
                if (s.run_crc != sg.want_crc || (uint32_t)s.run_len != sg.want_isize) { s.failed = true; return -1; }
                s.run_crc = 0;
                s.run_len = 0;
                s.in_member_data = false;
            }
        }
        lk.lock();
        s.consumed++;
        lk.unlock();
        s.cv.notify_all();
        if (pc.n_bytes == 0) {
            pc.drop_bytes();
            continue;
        }
        out = pc.bytes;
        len = pc.n_bytes;
        const size_t cap = pc.cap_bytes;
        keep = std::shared_ptr<void>(pc.bytes, [cap](void *p) { g_pool.put(p, cap); });
        pc.bytes = nullptr;
        pc.cap_bytes = 0;
        s.delivered += len;
        return 1;
    }
}

} // namespace kidhost
