// kid_pack_host.cpp - kid_pack_reads: the host-side writer of packed read batches (include/kmer_id.h).
//
// A FASTQ parser has every byte of a record in cache when it finds the line ends; turning the record
// into what the GPU scan consumes right there (process_qual's trim, newkmer_10nx.cpp:714-760, then
// 2 bits per surviving base) costs the parser a few instructions per base and cuts what crosses PCIe
// from 2 bytes per base (+ offsets) to 1/4 byte.  Qualities never leave the host.
//
// Three base packers with identical results: AVX2 (32 bases per step), BMI2-free 64-bit SWAR (8 bases
// per step) and the byte loop they are checked against in tests/test_pack_host_cpu.py.
#include "../../include/kmer_id.h"

#include <cstdlib>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>
#define KID_X86 1
#else
#define KID_X86 0
#endif

namespace {

// process_qual :724-753 on the bytes q[0..len): the (start, stop) it hands to process_read
inline void trim_span(const signed char *q, int len, int &start, int &stop)
{
    start = 0;
    stop = len - 1;
    if (len <= 0) return;
    while (start < stop && q[start] < 49) start++;      // :727-728  ('1' = 32 + 17)
    while (stop > start && q[stop] < 49) stop--;        // :729-730
    if (start < stop - 4) {                             // :732-742  leading 4-base window
        int w = q[start] + q[start + 1] + q[start + 2] + q[start + 3] - 128;
        while (w < 68 && start < stop - 4) {
            w += q[start + 4] - q[start];
            start++;
        }
    }
    if (start < stop - 4) {                             // :743-753  trailing window
        int w = q[stop] + q[stop - 1] + q[stop - 2] + q[stop - 3] - 128;
        while (w < 68 && start < stop - 4) {
            w += q[stop - 4] - q[stop];
            stop--;
        }
    }
}

struct Packed32 { // 32 bases: two code words and one validity word
    uint32_t c0, c1, valid;
};

// reference packer: one byte at a time (the definition of the format)
inline Packed32 pack32_bytes(const uint8_t *b, bool accept_u)
{
    Packed32 r = { 0, 0, 0 };
    for (int i = 0; i < 32; i++) {
        uint32_t code = 0, ok = 1;
        switch (b[i]) {
        case 'A': case 'a': code = 0; break;
        case 'C': case 'c': code = 1; break;
        case 'G': case 'g': code = 2; break;
        case 'T': case 't': code = 3; break;
        case 'U': case 'u': code = 3; ok = accept_u; break;
        default: ok = 0; break;
        }
        if (!ok) code = 0;
        if (i < 16) r.c0 |= code << (30 - 2 * i);
        else r.c1 |= code << (30 - 2 * (i - 16));
        r.valid |= ok << (31 - i);
    }
    return r;
}

// 64-bit SWAR: 8 bases per step.  Bits 2..1 of a letter are a raw code (A 0, C 1, T/U 2, G 3); ignoring
// those and the case bit an A/C/G byte equals 0x41 and a T byte 0x41 ^ 0x11 (U: ^ 0x10 more).
inline void pack8_swar(uint64_t x, bool accept_u, uint32_t &code16, uint32_t &valid8)
{
    const uint64_t L = 0x0101010101010101ull;
    const uint64_t s1 = x >> 1, s2 = x >> 2;
    const uint64_t tflag = s2 & ~s1 & L; // raw code 2
    uint64_t z = ((x ^ (0x41 * L)) & (0xD9 * L)) ^ (tflag * 0x11);
    if (accept_u) z &= ~tflag; // 'U' differs from 'T' in bit 0 only
    // 0x80 per NON-zero byte, exactly (no carries between bytes)
    const uint64_t nz = (((z & (0x7F * L)) + (0x7F * L)) | z) & (0x80 * L);
    const uint64_t ok = (nz ^ (0x80 * L)) >> 7; // 1 per accepted byte
    uint64_t c = s1 & (0x03 * L);
    c ^= (c >> 1) & L;       // swap 2 <-> 3: A0 C1 G2 T3
    c &= ok * 3;             // other bytes: code 0
    // gather: the first base (lowest byte) goes to the top
    const uint32_t lo = (uint32_t)c, hi = (uint32_t)(c >> 32);
    const uint32_t g0 = ((lo * 0x40100401u) >> 24) & 0xFFu, g1 = ((hi * 0x40100401u) >> 24) & 0xFFu;
    code16 = (g0 << 8) | g1;
    const uint32_t ol = (uint32_t)ok, oh = (uint32_t)(ok >> 32);
    const uint32_t v0 = ((ol * 0x08040201u) >> 24) & 0xFu, v1 = ((oh * 0x08040201u) >> 24) & 0xFu;
    valid8 = (v0 << 4) | v1;
}

inline Packed32 pack32_swar(const uint8_t *b, bool accept_u)
{
    Packed32 r;
    uint64_t x[4];
    memcpy(x, b, 32);
    uint32_t c[4], v[4];
    for (int i = 0; i < 4; i++) pack8_swar(x[i], accept_u, c[i], v[i]);
    r.c0 = (c[0] << 16) | c[1];
    r.c1 = (c[2] << 16) | c[3];
    r.valid = (v[0] << 24) | (v[1] << 16) | (v[2] << 8) | v[3];
    return r;
}

#if KID_X86
__attribute__((target("avx2"))) inline Packed32 pack32_avx2(const uint8_t *b, bool accept_u)
{
    const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(b));
    const __m256i up = _mm256_and_si256(x, _mm256_set1_epi8((char)0xDF)); // fold the case bit
    __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(up, _mm256_set1_epi8('A')),
                                                 _mm256_cmpeq_epi8(up, _mm256_set1_epi8('C'))),
                                 _mm256_or_si256(_mm256_cmpeq_epi8(up, _mm256_set1_epi8('G')),
                                                 _mm256_cmpeq_epi8(up, _mm256_set1_epi8('T'))));
    if (accept_u) ok = _mm256_or_si256(ok, _mm256_cmpeq_epi8(up, _mm256_set1_epi8('U')));
    __m256i c = _mm256_and_si256(_mm256_srli_epi16(x, 1), _mm256_set1_epi8(3)); // raw: A0 C1 T/U2 G3
    c = _mm256_xor_si256(c, _mm256_and_si256(_mm256_srli_epi16(c, 1), _mm256_set1_epi8(1)));
    c = _mm256_and_si256(c, ok);
    // 4 bases -> one byte, first base on top: (b0*4 + b1)*16 + (b2*4 + b3)
    const __m256i p2 = _mm256_maddubs_epi16(c, _mm256_set1_epi16(0x0104));
    const __m256i p4 = _mm256_madd_epi16(p2, _mm256_set1_epi32(0x00010010));
    // the low byte of each dword, last dword first -> one big-endian word per 128-bit lane
    const __m256i pick = _mm256_setr_epi8(12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                          12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i w = _mm256_shuffle_epi8(p4, pick);
    Packed32 r;
    r.c0 = (uint32_t)_mm256_extract_epi32(w, 0);
    r.c1 = (uint32_t)_mm256_extract_epi32(w, 4);
    // validity: byte 0 -> bit 31: reverse the 32 bytes, then movemask
    const __m256i rev = _mm256_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0,
                                         15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    const __m256i sh = _mm256_shuffle_epi8(ok, rev);           // reversed inside each 128-bit lane
    const __m256i sw = _mm256_permute2x128_si256(sh, sh, 0x01); // lanes swapped
    r.valid = (uint32_t)_mm256_movemask_epi8(sw);
    return r;
}
#endif

typedef Packed32 (*pack32_fn)(const uint8_t *, bool);

pack32_fn choose_packer(unsigned flags)
{
    if (flags & KID_PACK_IMPL_BYTES) return pack32_bytes; // tests: force an implementation
    if (flags & KID_PACK_IMPL_SWAR) return pack32_swar;
#if KID_X86
    if (__builtin_cpu_supports("avx2")) return pack32_avx2;
#endif
    return pack32_swar;
}

} // namespace

extern "C" {

size_t kid_pack_bound(size_t n_reads, uint64_t bases)
{
    return (size_t)(bases / 16 + bases / 32) + 2 * n_reads + 2;
}

int kid_pack_reads(const uint8_t *seq, const uint8_t *qual, const uint64_t *off, size_t n_reads, unsigned flags,
                   uint32_t word0, uint32_t *words, size_t words_cap, uint32_t *meta, uint32_t *span,
                   size_t *n_words)
{
    if (!meta || !n_words || (n_reads && (!seq || !off || !words))) return KID_EINVAL;
    const pack32_fn pack32 = choose_packer(flags);
    const bool accept_u = (flags & KID_DB_ACCEPT_U) != 0;
    size_t w = 0; // words written so far
    for (size_t r = 0; r < n_reads; r++) {
        const uint64_t o = off[r];
        const uint64_t len64 = off[r + 1] - o;
        if (len64 > 0x7FFFFFFFull) return KID_EINVAL; // one read < 2^31 bases
        const int len = (int)len64;
        int start = 0, stop = len - 1;
        if (qual) trim_span(reinterpret_cast<const signed char *>(qual) + o, len, start, stop);
        if (span) { span[2 * r] = (uint32_t)start; span[2 * r + 1] = (uint32_t)stop; }
        const int tlen = stop - start + 1;
        if ((uint64_t)word0 + w >= 0x80000000ull) return KID_ERANGE;
        meta[2 * r] = word0 + (uint32_t)w;
        meta[2 * r + 1] = 0;
        if (tlen <= KID_KSIZE) continue; // :755 - the read vanishes
        const size_t cw = ((size_t)tlen + 15) >> 4, vw = ((size_t)tlen + 31) >> 5;
        if (w + cw + vw > words_cap) return KID_ENOMEM;
        uint32_t *codes = words + w, *valid = codes + cw; // validity is kept only if some bit is clear
        const uint8_t *b = seq + o + (uint64_t)start;
        bool all_ok = true;
        size_t k = 0; // 32-base step
        for (; 32 * (k + 1) <= (size_t)tlen; k++) {
            const Packed32 pk = pack32(b + 32 * k, accept_u);
            codes[2 * k] = pk.c0;
            codes[2 * k + 1] = pk.c1;
            valid[k] = pk.valid;
            all_ok &= pk.valid == 0xFFFFFFFFu;
        }
        const int rem = tlen - (int)(32 * k);
        if (rem > 0) { // the tail goes through a padded copy: nothing is read past the read's last base
            uint8_t tmp[32];
            memset(tmp, 0, sizeof tmp);
            memcpy(tmp, b + 32 * k, (size_t)rem);
            const Packed32 pk = pack32(tmp, accept_u);
            codes[2 * k] = pk.c0;
            if (rem > 16) codes[2 * k + 1] = pk.c1;
            const uint32_t want = ~0u << (32 - rem);
            valid[k] = pk.valid; // bits beyond rem are 0: the padding is not a letter
            all_ok &= pk.valid == want;
        }
        meta[2 * r + 1] = (uint32_t)tlen;
        if (!all_ok) {
            meta[2 * r] |= KID_PK_INVALID;
            w += cw + vw;
        } else {
            w += cw;
        }
    }
    if ((uint64_t)word0 + w >= 0x80000000ull) return KID_ERANGE;
    meta[2 * n_reads] = word0 + (uint32_t)w;
    meta[2 * n_reads + 1] = 0;
    *n_words = w;
    return KID_OK;
}

size_t kid_dense_bound(uint64_t bases) { return (size_t)((bases + 15) / 16) + 4; }

int kid_pack_reads_dense(const uint8_t *seq, const uint8_t *qual, const uint64_t *off, size_t n_reads, unsigned flags,
                         uint32_t base0, uint32_t *codes, size_t codes_cap, uint32_t *boff, uint32_t *flagbits,
                         size_t read0, uint32_t *inv, size_t inv_cap, size_t *n_inv, uint32_t *span, uint32_t *n_bases)
{
    if (!boff || !flagbits || !n_inv || !n_bases || (n_reads && (!seq || !off || !codes))) return KID_EINVAL;
    const pack32_fn pack32 = choose_packer(flags);
    const bool accept_u = (flags & KID_DB_ACCEPT_U) != 0;
    // `codes` holds the stream from the 16-base unit that contains base0 on: word i = bases 16*(base0/16 + i) ...
    const uint64_t origin = (uint64_t)(base0 >> 4) << 4;
    uint64_t b = base0; // next free base
    size_t ni = 0;
    if ((b & 15) == 0 && codes_cap) codes[(b - origin) >> 4] = 0; // a word that is only partly ours is OR-ed into
    for (size_t r = 0; r < n_reads; r++) {
        const uint64_t o = off[r];
        const uint64_t len64 = off[r + 1] - o;
        if (len64 > 0x7FFFFFFFull) return KID_EINVAL; // one read < 2^31 bases
        const int len = (int)len64;
        int start = 0, stop = len - 1;
        if (qual) trim_span(reinterpret_cast<const signed char *>(qual) + o, len, start, stop);
        if (span) { span[2 * r] = (uint32_t)start; span[2 * r + 1] = (uint32_t)stop; }
        const int tlen = stop - start + 1;
        const size_t gr = read0 + r; // position of this read's flag bit
        if ((gr & 31) == 0) flagbits[gr >> 5] = 0;
        boff[r] = (uint32_t)b;
        if (tlen <= KID_KSIZE) continue; // :755 - the read vanishes: no bases
        if (b + (uint64_t)tlen >= 0xFFFFFFF0ull) return KID_ERANGE; // base offsets are 32-bit: split the batch
        if (((b + (uint64_t)tlen - origin + 15) >> 4) + 3 > codes_cap) return KID_ENOMEM;
        const uint8_t *src = seq + o + (uint64_t)start;
        bool flagged = false;
        for (int k = 0; k < tlen; k += 32) {
            const int rem = tlen - k < 32 ? tlen - k : 32;
            Packed32 pk;
            if (rem == 32) {
                pk = pack32(src + k, accept_u);
            } else { // the tail goes through a padded copy: nothing is read past the read's last base
                uint8_t tmp[32];
                memset(tmp, 0, sizeof tmp);
                memcpy(tmp, src + k, (size_t)rem);
                pk = pack32(tmp, accept_u);
            }
            // append 64 bits (32 bases; those past rem are 0) at base b: bit 2*(b%16) of word (b - origin)/16, MSB first
            const uint64_t v = ((uint64_t)pk.c0 << 32) | pk.c1;
            const size_t w = (size_t)((b - origin) >> 4);
            const unsigned sh = (unsigned)(b & 15) * 2;
            if (sh == 0) {
                codes[w] = (uint32_t)(v >> 32);
                codes[w + 1] = (uint32_t)v;
                codes[w + 2] = 0;
            } else {
                codes[w] |= (uint32_t)(v >> (32 + sh));
                codes[w + 1] = (uint32_t)(v >> sh);
                codes[w + 2] = (uint32_t)(v << (32 - sh));
            }
            const uint32_t want = rem == 32 ? 0xFFFFFFFFu : ~0u << (32 - rem);
            uint32_t bad = ~pk.valid & want;
            if (bad) {
                flagged = true;
                while (bad) { // positions of the bases that are not ACGT(+U), ascending
                    const int i = __builtin_clz(bad);
                    bad &= ~(0x80000000u >> i);
                    if (ni >= inv_cap) return KID_ENOMEM;
                    inv[ni++] = (uint32_t)(b + (uint64_t)i);
                }
            }
            b += (uint64_t)rem;
        }
        if (flagged) flagbits[gr >> 5] |= 1u << (gr & 31);
    }
    boff[n_reads] = (uint32_t)b;
    *n_inv = ni;
    *n_bases = (uint32_t)b;
    return KID_OK;
}

} // extern "C"
