// kid_inflate.cuh - deflate (RFC 1951) decoding for the device-side gzip FASTQ reader (kid_ingest.cu).
//
// Replaces the gzread loop of process_fqgz (newkmer_10nx.cpp:770-780) for whole files: the compressed
// bytes cross PCIe and are inflated on the GPU.  A gzip member cannot be split up front (a block may
// start at any bit and refers to the 32 KiB before it), so the file is cut into fixed-size pieces and
// each piece is inflated speculatively, the scheme of host/pgz.cpp laid out for the GPU:
//   find    one warp per piece scans for the first bit that parses as a dynamic-Huffman block header
//           with complete codes (is_block_start_candidate + parse_dynamic_header);
//   inflate one THREAD per piece (Huffman decoding is serial) decodes from there to the first block
//           boundary at or after the end of its piece into 16-bit symbols: a byte, or "copy from dist
//           back" for every position of a match (no loads: the decoder never waits for memory);
//   copy    one WARP per piece turns the copies into symbols, 32 positions at a time: a byte, or a
//           marker "byte i of the 32 KiB before this piece" (256 + i);
//   chain   (host) a piece counts only if it started exactly where its predecessor stopped;
//   resolve markers are replaced through per-piece window maps (kid_ingest.cu).
// Everything here is __host__ __device__ and free of CUDA intrinsics so that the CPU tests
// (tests/hosttest/inflate_emul.cpp) run the very same code against zlib.
//
// The decode tables of one decoder are 16-bit entries addressed as base[i * STRIDE]: on the device the
// 32 decoders of a warp interleave their tables in shared memory (STRIDE = 32, at most two-way bank
// conflicts), on the host STRIDE = 1.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KIDZ_HD __host__ __device__ __forceinline__
#else
#define KIDZ_HD inline
#endif

// a refusal inside inflate_piece; the CPU harness can make it say where (-DKIDZ_TRACE)
#if defined(KIDZ_TRACE) && !defined(__CUDA_ARCH__)
#include <stdio.h>
#define KIDZ_REFUSE do { fprintf(stderr, "inflate_piece: refused at line %d (status %u)\n", __LINE__, res.status); return; } while (0)
#else
#define KIDZ_REFUSE return
#endif

namespace kidz {

constexpr int kWin = 32768;  // deflate history
constexpr int kLitRoot = 9;  // bits of the literal/length fast table
constexpr int kDistRoot = 8; // bits of the distance fast table
constexpr int kClRoot = 7;   // code-length code: every code fits the fast table

// one decoder's table area, in 16-bit entries
constexpr int kOffLitFast = 0;                             // (symbol << 4) | code bits, 0 = longer than kLitRoot
constexpr int kOffDistFast = kOffLitFast + (1 << kLitRoot);
constexpr int kOffLitSym = kOffDistFast + (1 << kDistRoot); // symbols in canonical order (the slow path)
constexpr int kOffDistSym = kOffLitSym + 288;
constexpr int kOffLitCnt = kOffDistSym + 32;               // codes per length 0..15
constexpr int kOffDistCnt = kOffLitCnt + 16;
constexpr int kTabEntries = kOffDistCnt + 16;              // 1120 entries = 2240 bytes
// the code-length code lives where the literal table will be built afterwards
constexpr int kOffClFast = kOffLitFast;                    // 128 entries
constexpr int kOffClCnt = kOffLitFast + 128;               // 16
constexpr int kOffClSym = kOffLitFast + 144;               // 19
// the block finder only needs the code-length code
constexpr int kFindTabEntries = 128 + 16 + 32;

constexpr int kMaxEnds = 6; // gzip members that may end inside one piece
// what inflate_piece writes for a position inside a match, until resolve_copies replaces it by the
// symbol `dist` positions back (distances are 1..32768)
constexpr uint32_t kCopyFlag = 0x8000u;

enum PieceStatus : uint32_t {
    kPieceOk = 0,
    kPieceNoStart = 1,  // no block start found inside the piece
    kPieceBadData = 2,  // not a deflate stream from this start (or a stream zlib has to judge)
    kPieceOverflow = 3, // more output than the piece's slot holds
    kPieceManyEnds = 4, // more than kMaxEnds members end in it
};

struct MemberEnd {
    uint32_t out_pos; // symbols of this piece that belong to the member ending here
    uint32_t crc, isize;
};

struct PieceResult {
    uint64_t start_bit, end_bit;
    uint32_t n_out;
    uint32_t status;
    uint32_t n_ends;
    uint32_t eof; // the stream ended cleanly inside this piece
    MemberEnd ends[kMaxEnds];
};

template <int STRIDE>
struct Tab {
    uint16_t *base;
    KIDZ_HD uint16_t &at(int i) const { return base[i * STRIDE]; }
};

// ---- bit reader, LSB first, over 32-bit words (the buffer is padded with >= 16 zero bytes) ----------
struct BitIn {
    const uint32_t *w;
    uint64_t buf;
    uint32_t cnt;
    uint64_t next; // next word to load
    KIDZ_HD void refill() // afterwards at least 32 bits are available
    {
        if (cnt <= 32) {
            buf |= (uint64_t)w[next++] << cnt;
            cnt += 32;
        }
    }
    KIDZ_HD void seek(const uint32_t *words, uint64_t bit)
    {
        w = words;
        next = bit >> 5;
        buf = 0;
        cnt = 0;
        refill();
        drop((uint32_t)(bit & 31));
        refill();
    }
    KIDZ_HD void drop(uint32_t n) { buf >>= n; cnt -= n; }
    KIDZ_HD uint32_t take(uint32_t n)
    {
        const uint32_t v = (uint32_t)buf & ((1u << n) - 1u);
        drop(n);
        return v;
    }
    KIDZ_HD uint64_t pos() const { return next * 32 - cnt; }
};

KIDZ_HD uint32_t reverse_bits(uint32_t code, int n) // the low n bits of code, mirrored
{
    uint32_t r = 0;
    for (int i = 0; i < n; i++) {
        r = (r << 1) | (code & 1u);
        code >>= 1;
    }
    return r;
}

enum CodeKind { kCodeLit, kCodeDist, kCodeCl };

// Canonical Huffman code from code lengths: count per length + symbols in canonical order + a fast
// table of `root` bits.  false = a code inflate() refuses (over-subscribed, or incomplete where it
// does not allow that) - or one this decoder is stricter about; every refusal ends in the host reader.
template <int STRIDE>
KIDZ_HD bool build_code(const uint8_t *lens, int n, int root, CodeKind kind, Tab<STRIDE> t, int off_fast, int off_cnt,
                        int off_sym)
{
    for (int l = 0; l < 16; l++) t.at(off_cnt + l) = 0;
    for (int s = 0; s < n; s++) t.at(off_cnt + lens[s])++;
    const int used = n - t.at(off_cnt);
    for (int i = 0; i < (1 << root); i++) t.at(off_fast + i) = 0;
    if (used == 0) {
        t.at(off_cnt) = 0;
        return kind == kCodeDist; // no distance codes: legal, any match is then an error
    }
    int left = 1, maxlen = 0;
    for (int l = 1; l <= 15; l++) {
        left = (left << 1) - (int)t.at(off_cnt + l);
        if (left < 0) return false;
        if (t.at(off_cnt + l)) maxlen = l;
    }
    if (left > 0 && !(kind == kCodeDist && used == 1 && maxlen == 1)) return false;
    t.at(off_cnt) = 0;
    uint16_t offs[16]; // first index of each length in the sorted symbols
    uint16_t next[16]; // first code of each length
    offs[1] = 0;
    next[1] = 0;
    for (int l = 1; l < 15; l++) {
        offs[l + 1] = (uint16_t)(offs[l] + t.at(off_cnt + l));
        next[l + 1] = (uint16_t)((next[l] + t.at(off_cnt + l)) << 1);
    }
    for (int s = 0; s < n; s++) {
        const int l = lens[s];
        if (!l) continue;
        t.at(off_sym + offs[l]++) = (uint16_t)s;
        const uint32_t code = next[l]++;
        if (l <= root) {
            const uint16_t e = (uint16_t)((s << 4) | l);
            for (uint32_t i = reverse_bits(code, l); i < (1u << root); i += 1u << l) t.at(off_fast + (int)i) = e;
        }
    }
    return true;
}

// one symbol of a code: fast table, else bit by bit in canonical order; -1 = no such code
template <int STRIDE>
KIDZ_HD int decode_symbol(BitIn &in, Tab<STRIDE> t, int root, int off_fast, int off_cnt, int off_sym)
{
    const uint32_t e = t.at(off_fast + (int)((uint32_t)in.buf & ((1u << root) - 1u)));
    if (e & 15u) {
        in.drop(e & 15u);
        return (int)(e >> 4);
    }
    int code = 0, first = 0, index = 0;
    for (int l = 1; l <= 15; l++) {
        code |= (int)((in.buf >> (l - 1)) & 1u);
        const int count = t.at(off_cnt + l);
        if (code - count < first) {
            in.drop((uint32_t)l);
            return t.at(off_sym + index + (code - first));
        }
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// Dynamic block header after the three block-type bits: code lengths of the literal/length and
// distance codes into lens[0..nlen) and lens[nlen..nlen+ndist).  Mirrors inflate's checks.
// The code-length code is built at (off_fast, off_cnt, off_sym) of t.
template <int STRIDE>
KIDZ_HD bool parse_dynamic_header(BitIn &in, Tab<STRIDE> t, int off_fast, int off_cnt, int off_sym, uint8_t *lens,
                                  int &nlen, int &ndist, bool text_only = false)
{
    const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
    in.refill();
    nlen = (int)in.take(5) + 257;
    ndist = (int)in.take(5) + 1;
    const int ncl = (int)in.take(4) + 4;
    if (nlen > 286 || ndist > 30) return false;
    uint8_t cl[19];
    for (int i = 0; i < 19; i++) cl[i] = 0;
    for (int i = 0; i < ncl; i++) {
        in.refill();
        cl[order[i]] = (uint8_t)in.take(3);
    }
    if (!build_code(cl, 19, kClRoot, kCodeCl, t, off_fast, off_cnt, off_sym)) return false;
    int have = 0;
    const int total = nlen + ndist;
    // code space used so far, in units of 2^-15: an over-subscribed code (what a bit position that is not a
    // block header nearly always yields) is refused as soon as it shows, not after all lengths are read
    uint32_t used_lit = 0, used_dist = 0;
    while (have < total) {
        in.refill();
        const int sym = decode_symbol(in, t, kClRoot, off_fast, off_cnt, off_sym);
        if (sym < 0) return false;
        if (sym < 16) {
            if (sym) {
                if (text_only && (have < 9 || (have >= 128 && have < 256))) return false; // see is_block_start
                if (have < nlen) used_lit += 32768u >> sym; else used_dist += 32768u >> sym;
                if (used_lit > 32768u || used_dist > 32768u) return false;
            }
            lens[have++] = (uint8_t)sym;
            continue;
        }
        int rep;
        uint8_t val = 0;
        if (sym == 16) {
            if (have == 0) return false;
            val = lens[have - 1];
            rep = 3 + (int)in.take(2);
        } else if (sym == 17) rep = 3 + (int)in.take(3);
        else rep = 11 + (int)in.take(7);
        if (have + rep > total) return false;
        if (text_only && val && (have < 9 || (have + rep > 128 && have < 256))) return false;
        for (int i = 0; i < rep; i++) {
            if (val) {
                if (have < nlen) used_lit += 32768u >> val; else used_dist += 32768u >> val;
            }
            lens[have++] = val;
        }
        if (used_lit > 32768u || used_dist > 32768u) return false;
    }
    return lens[256] != 0; // a block needs its end-of-block code
}

// Complete-code test of build_code without building anything (the block finder).
KIDZ_HD bool code_is_acceptable(const uint8_t *lens, int n, CodeKind kind)
{
    int count[16];
    for (int l = 0; l < 16; l++) count[l] = 0;
    for (int s = 0; s < n; s++) count[lens[s]]++;
    const int used = n - count[0];
    if (used == 0) return kind == kCodeDist;
    int left = 1, maxlen = 0;
    for (int l = 1; l <= 15; l++) {
        left = (left << 1) - count[l];
        if (left < 0) return false;
        if (count[l]) maxlen = l;
    }
    return !(left > 0 && !(kind == kCodeDist && used == 1 && maxlen == 1));
}

KIDZ_HD uint32_t peek32(const uint32_t *w, uint32_t bit) // 32 bits from any bit position below 2^32
{
    const uint32_t i = bit >> 5, sh = bit & 31u;
    const uint64_t v = (uint64_t)w[i] | ((uint64_t)w[i + 1] << 32);
    return (uint32_t)(v >> sh);
}
KIDZ_HD uint32_t peek32(const uint32_t *w, uint64_t bit)
{
    const uint64_t i = bit >> 5;
    const uint32_t sh = (uint32_t)(bit & 31);
    const uint64_t v = (uint64_t)w[i] | ((uint64_t)w[i + 1] << 32);
    return (uint32_t)(v >> sh);
}

// Cheap tests of a bit position.  First: BTYPE = 2 and plausible counts (v = the 32 bits at `bit`)...
KIDZ_HD bool block_start_bits_plausible(uint32_t v)
{
    return (v & 6u) == 4u && ((v >> 3) & 31u) <= 29u && ((v >> 8) & 31u) <= 29u;
}
// ...second: a COMPLETE code-length code.  (BitPos: uint64_t, or uint32_t for positions relative to a
// piece - the block finder's inner loops stay in 32-bit arithmetic.)
template <class BitPos>
KIDZ_HD bool block_start_cl_complete(const uint32_t *w, BitPos bit, uint32_t v)
{
    const int ncl = (int)((v >> 13) & 15u) + 4;
    uint32_t kraft = 0;
    const uint32_t a = peek32(w, bit + 17), b = peek32(w, bit + 47);
    for (int i = 0; i < 19; i++) {
        const uint32_t l = (i < 10 ? a >> (3 * i) : b >> (3 * (i - 10))) & 7u;
        if (i < ncl && l) kraft += 128u >> l;
    }
    return kraft == 128u;
}
KIDZ_HD bool is_block_start_candidate(const uint32_t *w, uint64_t bit)
{
    const uint32_t v = peek32(w, bit);
    return block_start_bits_plausible(v) && block_start_cl_complete(w, bit, v);
}

// Full test: the whole dynamic header parses and both codes are ones build_code accepts.
// text_only: also require that no byte below 0x09 or from 0x80 on has a code.  This is a filter on
// SPECULATION only (a real block that fails it is found again by the chain walk, from its predecessor's
// end): about one bit position in 4e8 passes the header test by chance, i.e. once per FASTQ file, and
// costs a piece that has to be inflated a second time, by one warp; a block of FASTQ text never codes such
// bytes, a chance header nearly always does - with its very first code length, which is also what makes
// this test cheap.
template <int STRIDE>
KIDZ_HD bool is_block_start(const uint32_t *w, uint64_t bit, Tab<STRIDE> t, bool text_only = false)
{
    BitIn in;
    in.seek(w, bit + 3);
    uint8_t lens[286 + 30 + 4];
    int nlen, ndist;
    if (!parse_dynamic_header(in, t, 0, 128, 144, lens, nlen, ndist, text_only)) return false;
    return code_is_acceptable(lens, nlen, kCodeLit) && code_is_acceptable(lens + nlen, ndist, kCodeDist);
}

// gzip member header at byte `at`: its length, 0 if it is not one this decoder accepts
KIDZ_HD uint64_t gzip_header_len(const uint8_t *d, uint64_t size, uint64_t at)
{
    if (at + 10 > size || d[at] != 0x1f || d[at + 1] != 0x8b || d[at + 2] != 8 || (d[at + 3] & 0xe0)) return 0;
    const uint8_t flg = d[at + 3];
    uint64_t p = at + 10;
    if (flg & 4) { // FEXTRA
        if (p + 2 > size) return 0;
        p += 2 + ((uint64_t)d[p] | ((uint64_t)d[p + 1] << 8));
    }
    for (int f = 8; f <= 16; f <<= 1) // FNAME, FCOMMENT: zero terminated
        if (flg & f) {
            while (p < size && d[p]) p++;
            p++;
        }
    if (flg & 2) p += 2; // FHCRC
    return p < size ? p - at : 0;
}

// Inflates from start_bit (a block header) up to the first block boundary at or after stop_bit, or the
// clean end of the file, into 16-bit symbols.  `floor0` = how far back a match may reach at the start:
// kWin inside a member of unknown history (markers), 0 at the first block of a member.
//
// A state machine advanced by step(): the 32 decoders of a warp run in lockstep on the device (a
// __syncwarp before every step), which only works if one step is short whatever the data: one block
// header, or one or two symbols, or 16 positions of a long match.  (Left to themselves the lanes of a warp
// drift apart in a data-dependent loop and the warp ends up executing them one at a time: measured, the
// free-running version of this decoder took 0.29 s for 3844 pieces.)
template <int STRIDE>
struct Inflater {
    enum State : uint32_t { kAtBoundary = 0, kInBlock = 1, kDone = 2 };
    BitIn in;
    const uint32_t *w;
    uint64_t size, stop_bit;
    uint16_t *out;
    int32_t cap, pos, floor; // symbols: room, written (incl. a fill in progress), lowest position a match may read
    Tab<STRIDE> t;
    PieceResult *res;
    uint32_t state, bfinal;
    uint32_t fill_left;      // positions of the current match still to be written
    int32_t fill_at;
    uint32_t fill_code;

    KIDZ_HD void start(const uint32_t *words, uint64_t size_, uint64_t start_bit, uint64_t stop_bit_, int floor0, uint16_t *out_,
                       uint32_t out_cap, Tab<STRIDE> t_, PieceResult *res_)
    {
        w = words;
        size = size_;
        stop_bit = stop_bit_;
        out = out_;
        cap = (int32_t)out_cap; // < 2^31
        pos = 0;
        floor = -floor0;
        t = t_;
        res = res_;
        state = kAtBoundary;
        bfinal = 0;
        fill_left = 0;
        fill_at = 0;
        fill_code = 0;
        res->start_bit = start_bit;
        res->end_bit = start_bit;
        res->n_out = 0;
        res->n_ends = 0;
        res->eof = 0;
        res->status = kPieceBadData;
        in.seek(w, start_bit);
    }
    KIDZ_HD bool done() const { return state == kDone; }
    KIDZ_HD void refuse(uint32_t status) { res->status = status; state = kDone; }
    KIDZ_HD void finish(uint64_t end_bit, uint32_t eof)
    {
        res->end_bit = end_bit;
        res->n_out = (uint32_t)pos;
        res->eof = eof;
        res->status = kPieceOk;
        state = kDone;
    }

    KIDZ_HD void write_fill() // up to 16 positions of "copy from dist back" (kCopyFlag | dist - 1)
    {
        const uint32_t n = fill_left < 16u ? fill_left : 16u;
        const uint16_t code = (uint16_t)fill_code;
        const uint32_t two = fill_code * 0x10001u;
        uint32_t i = 0;
        if (fill_at & 1) out[fill_at] = code, i = 1;
        for (; i + 1 < n; i += 2) *reinterpret_cast<uint32_t *>(out + fill_at + (int32_t)i) = two;
        if (i < n) out[fill_at + (int32_t)i] = code;
        fill_at += (int32_t)n;
        fill_left -= n;
    }

    KIDZ_HD void end_of_block()
    {
        const uint8_t *bytes = reinterpret_cast<const uint8_t *>(w);
        const uint64_t size_bits = size * 8;
        if (in.pos() > size_bits) return refuse(kPieceBadData);
        state = kAtBoundary;
        if (!bfinal) return;
        // member trailer, then the next member or the end of the file
        uint64_t at = (in.pos() + 7) >> 3;
        if (at + 8 > size) return refuse(kPieceBadData);
        if (res->n_ends >= (uint32_t)kMaxEnds) return refuse(kPieceManyEnds);
        MemberEnd &me = res->ends[res->n_ends++];
        me.out_pos = (uint32_t)pos;
        me.crc = (uint32_t)bytes[at] | ((uint32_t)bytes[at + 1] << 8) | ((uint32_t)bytes[at + 2] << 16) | ((uint32_t)bytes[at + 3] << 24);
        me.isize = (uint32_t)bytes[at + 4] | ((uint32_t)bytes[at + 5] << 8) | ((uint32_t)bytes[at + 6] << 16) | ((uint32_t)bytes[at + 7] << 24);
        at += 8;
        if (at == size) return finish(size_bits, 1);
        const uint64_t hl = gzip_header_len(bytes, size, at);
        if (!hl) return refuse(kPieceBadData); // trailing bytes that are not another member: zlib's business
        floor = pos;
        in.seek(w, (at + hl) * 8);
    }

    KIDZ_HD void block_header()
    {
        const uint8_t *bytes = reinterpret_cast<const uint8_t *>(w);
        const uint64_t size_bits = size * 8;
        const uint64_t here = in.pos();
        if (here + 3 > size_bits) return refuse(kPieceBadData);
        in.refill();
        // a piece ends at the first block boundary at or after stop_bit whose block has dynamic codes: the
        // only kind of start the block finder can see (stored and fixed blocks there are inflated as well)
        if (here >= stop_bit && ((uint32_t)in.buf & 6u) == 4u) return finish(here, 0);
        bfinal = in.take(1);
        const uint32_t btype = in.take(2);
        if (btype == 3) return refuse(kPieceBadData);
        if (btype == 0) { // stored
            uint64_t at = (in.pos() + 7) >> 3;
            if (at + 4 > size) return refuse(kPieceBadData);
            const uint32_t len = bytes[at] | ((uint32_t)bytes[at + 1] << 8), nlen = bytes[at + 2] | ((uint32_t)bytes[at + 3] << 8);
            if ((len ^ 0xffffu) != nlen) return refuse(kPieceBadData);
            at += 4;
            if (at + len > size) return refuse(kPieceBadData);
            if (len > (uint32_t)(cap - pos)) return refuse(kPieceOverflow);
            for (uint32_t i = 0; i < len; i++) out[pos + (int32_t)i] = bytes[at + i];
            pos += (int32_t)len;
            in.seek(w, (at + len) * 8);
            return end_of_block();
        }
        if (btype == 2) {
            uint8_t lens[286 + 30 + 4];
            int nlen, ndist;
            if (!parse_dynamic_header(in, t, kOffClFast, kOffClCnt, kOffClSym, lens, nlen, ndist)) return refuse(kPieceBadData);
            if (!build_code(lens, nlen, kLitRoot, kCodeLit, t, kOffLitFast, kOffLitCnt, kOffLitSym)) return refuse(kPieceBadData);
            if (!build_code(lens + nlen, ndist, kDistRoot, kCodeDist, t, kOffDistFast, kOffDistCnt, kOffDistSym)) return refuse(kPieceBadData);
        } else {
            uint8_t lens[288];
            for (int s = 0; s < 288; s++) lens[s] = (uint8_t)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8);
            build_code(lens, 288, kLitRoot, kCodeLit, t, kOffLitFast, kOffLitCnt, kOffLitSym);
            for (int s = 0; s < 32; s++) lens[s] = 5; // 30 and 31 never occur in valid data
            build_code(lens, 32, kDistRoot, kCodeDist, t, kOffDistFast, kOffDistCnt, kOffDistSym);
        }
        state = kInBlock;
    }

    KIDZ_HD void symbol()
    {
        in.refill();
        if (in.next * 4 > size + 16) return refuse(kPieceBadData); // ran off the end of the file
        const int sym = decode_symbol(in, t, kLitRoot, kOffLitFast, kOffLitCnt, kOffLitSym);
        if (sym < 0) return refuse(kPieceBadData);
        if (sym < 256) {
            if (pos + 2 > cap) return refuse(kPieceOverflow);
            out[pos++] = (uint16_t)sym;
            // a second literal from the same refill (>= 17 bits are left: enough for any code)
            const uint32_t e = t.at(kOffLitFast + (int)((uint32_t)in.buf & ((1u << kLitRoot) - 1u)));
            if ((e & 15u) && (e >> 4) < 256u) {
                in.drop(e & 15u);
                out[pos++] = (uint16_t)(e >> 4);
            }
            return;
        }
        if (sym == 256) return end_of_block();
        if (sym > 285) return refuse(kPieceBadData);
        uint32_t len;
        if (sym < 265) len = (uint32_t)sym - 254u;
        else if (sym == 285) len = 258;
        else {
            const uint32_t x = (uint32_t)sym - 261u, eb = x >> 2;
            len = 3u + ((4u + (x & 3u)) << eb) + in.take(eb);
        }
        in.refill();
        const int ds = decode_symbol(in, t, kDistRoot, kOffDistFast, kOffDistCnt, kOffDistSym);
        if (ds < 0 || ds > 29) return refuse(kPieceBadData);
        uint32_t dist;
        if (ds < 4) dist = (uint32_t)ds + 1u;
        else {
            const uint32_t eb = ((uint32_t)ds >> 1) - 1u;
            dist = 1u + ((2u + ((uint32_t)ds & 1u)) << eb) + in.take(eb);
        }
        if ((int32_t)dist > pos - floor) return refuse(kPieceBadData); // before the member / before any history
        if ((int32_t)len > cap - pos) return refuse(kPieceOverflow);
        // the copy itself is left to resolve_copies: every position of the match gets "copy from dist
        // back", so that this decoder never waits for memory
        fill_code = kCopyFlag | (dist - 1u);
        fill_at = pos;
        fill_left = len;
        pos += (int32_t)len;
        write_fill();
    }

    KIDZ_HD void step()
    {
        if (state == kInBlock) {
            if (fill_left) write_fill();
            else symbol();
        } else if (state == kAtBoundary) block_header();
    }
};

template <int STRIDE>
KIDZ_HD void inflate_piece(const uint32_t *w, uint64_t size, uint64_t start_bit, uint64_t stop_bit, int floor0,
                           uint16_t *out, uint32_t out_cap, Tab<STRIDE> t, PieceResult &res)
{
    Inflater<STRIDE> d;
    d.start(w, size, start_bit, stop_bit, floor0, out, out_cap, t, &res);
    while (!d.done()) d.step();
}

// Second pass over a piece's output, in order: every "copy from dist back" becomes the symbol it refers
// to - an earlier symbol of the piece, or a marker for the unknown 32 KiB before it.  This is the plain
// definition; the device runs it with a warp per piece, 32 positions at a time (kid_ingest.cu).
KIDZ_HD void resolve_copies(uint16_t *out, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t x = out[i];
        if (!(x & kCopyFlag)) continue;
        const int32_t src = (int32_t)i - (int32_t)((x & 0x7fffu) + 1u);
        out[i] = src >= 0 ? out[src] : (uint16_t)(256 + kWin + src);
    }
}

// ---- CRC-32 (the gzip check value) in pieces ---------------------------------------------------------
// crc(A ++ B) = crc(A) * x^(8 |B|) mod P  xor  crc(B) over GF(2), with the reflected polynomial
// 0xedb88320 (RFC 1952 section 8): every chunk's CRC is shifted by the bytes that follow it in its
// member and the shifted values are xor-ed together.
constexpr uint32_t kCrcPoly = 0xedb88320u;

KIDZ_HD uint32_t crc_mulmod(uint32_t a, uint32_t b) // a(x) * b(x) mod P, reflected representation
{
    uint32_t p = 0;
    for (uint32_t m = 0x80000000u; m; m >>= 1) {
        if (a & m) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ kCrcPoly : b >> 1;
    }
    return p;
}

// x^(8 n) mod P; pow2[k] = x^(2^k) mod P
KIDZ_HD uint32_t crc_x8n(uint64_t n, const uint32_t *pow2)
{
    uint32_t p = 0x80000000u; // the polynomial 1
    int k = 3;                // x^(8 n) = x^(n * 2^3)
    while (n) {
        if (n & 1) p = crc_mulmod(pow2[k & 31], p);
        n >>= 1;
        k++;
    }
    return p;
}

inline void crc_make_pow2(uint32_t *pow2) // host side, once
{
    uint32_t p = 0x40000000u; // x^1
    pow2[0] = p;
    for (int k = 1; k < 32; k++) pow2[k] = p = crc_mulmod(p, p);
}

inline void crc_make_table(uint32_t *tab) // host side, once
{
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ kCrcPoly : c >> 1;
        tab[i] = c;
    }
}

} // namespace kidz
