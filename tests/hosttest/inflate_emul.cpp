// inflate_emul.cpp - runs the device-side inflate of kmer_id_b200/csrc/kid_inflate.cuh on the CPU, step
// by step the way kid_ingest.cu runs it on the GPU (find -> inflate per piece -> chain -> window maps in
// groups -> resolve -> CRC-32 in chunks), and compares the text with zlib's.  Test infrastructure.
//   inflate_emul FILE.gz [piece_bytes [expand [group]]]      exit 0 = identical, 3 = cleanly refused
//   KIDZ_TEXT_ONLY=1: the block finder's text-only filter; KIDZ_WARP=1: the warp-per-piece decoder (emulated)
#include "../../kmer_id_b200/csrc/kid_inflate_chain.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <zlib.h>

using namespace kidz;

// ---- the warp-per-piece decoder of kid_ingest.cu (warp_symbols / warp_inflate_piece), its 32 lanes played one
// after the other: every lane decodes the token that would start at "its" bit, the real tokens are the chain
// from lane 0, block headers and tokens the fast tables do not cover go through the plain decoder.  A
// restatement for the CPU (the device code uses shuffles and cannot be compiled here), statement by statement.
enum { kRoundEob = 0, kRoundBad = 1, kRoundOverflow = 2, kRoundSlowToken = 3 };

static int warp_symbols_emul(const uint32_t *w, uint64_t size_bits, Tab<1> t, uint16_t *out, uint64_t &bp, int32_t &pos, int32_t floor,
                             int32_t cap)
{
    for (;;) {
        if (bp > size_bits) return kRoundBad;
        uint32_t bits[32], outlen[32], code[32], kind[32];
        for (uint32_t lane = 0; lane < 32; lane++) {
            const uint64_t b = bp + lane;
            const uint32_t lo = peek32(w, b), hi = peek32(w, b + 32);
            const uint64_t v = (uint64_t)lo | ((uint64_t)hi << 32);
            const uint32_t e = t.at(kOffLitFast + (int)(lo & ((1u << kLitRoot) - 1u)));
            bits[lane] = e & 15u;
            outlen[lane] = 1;
            code[lane] = e >> 4;
            kind[lane] = 0;
            if (bits[lane] == 0 || code[lane] > 285u) kind[lane] = 3;
            else if (code[lane] == 256u) { kind[lane] = 2; outlen[lane] = 0; }
            else if (code[lane] > 256u) {
                uint32_t len;
                const uint32_t c = code[lane];
                if (c < 265u) len = c - 254u;
                else if (c == 285u) len = 258u;
                else {
                    const uint32_t x = c - 261u, eb = x >> 2;
                    len = 3u + ((4u + (x & 3u)) << eb) + ((uint32_t)(v >> bits[lane]) & ((1u << eb) - 1u));
                    bits[lane] += eb;
                }
                const uint32_t de = t.at(kOffDistFast + (int)((uint32_t)(v >> bits[lane]) & ((1u << kDistRoot) - 1u)));
                const uint32_t ds = de >> 4;
                if ((de & 15u) == 0 || ds > 29u) kind[lane] = 3;
                else {
                    bits[lane] += de & 15u;
                    uint32_t dist;
                    if (ds < 4u) dist = ds + 1u;
                    else {
                        const uint32_t eb = (ds >> 1) - 1u;
                        dist = 1u + ((2u + (ds & 1u)) << eb) + ((uint32_t)(v >> bits[lane]) & ((1u << eb) - 1u));
                        bits[lane] += eb;
                    }
                    kind[lane] = 1;
                    outlen[lane] = len;
                    code[lane] = kCopyFlag | (dist - 1u);
                }
            }
        }
        uint32_t cur = 0, total = 0;
        int32_t at[32];
        bool on[32] = { false }, eob = false;
        while (cur < 32u) {
            const uint32_t kd = kind[cur];
            if (kd == 3u) break;
            on[cur] = true;
            at[cur] = pos + (int32_t)total;
            total += outlen[cur];
            cur = cur + bits[cur];
            if (kd == 2u) { eob = true; break; }
        }
        if (total == 0 && !eob && cur == 0) return kRoundSlowToken;
        if (total > (uint32_t)(cap - pos)) return kRoundOverflow;
        for (uint32_t lane = 0; lane < 32; lane++)
            if (on[lane] && kind[lane] == 1u && (int32_t)((code[lane] & 0x7fffu) + 1u) > at[lane] - floor) return kRoundBad;
        for (uint32_t lane = 0; lane < 32; lane++) {
            if (!on[lane]) continue;
            if (kind[lane] == 0u) out[at[lane]] = (uint16_t)code[lane];
            if (kind[lane] == 1u)
                for (uint32_t i = 0; i < outlen[lane]; i++) out[at[lane] + (int32_t)i] = (uint16_t)code[lane];
        }
        pos += (int32_t)total;
        bp += cur;
        if (eob) return kRoundEob;
    }
}

static void warp_inflate_piece_emul(const uint32_t *w, uint64_t size, uint64_t start_bit, uint64_t stop_bit, int floor0, uint16_t *out,
                                    uint32_t out_cap, Tab<1> tab, PieceResult &res)
{
    Inflater<1> d;
    d.start(w, size, start_bit, stop_bit, floor0, out, out_cap, tab, &res);
    for (;;) {
        if (d.state == Inflater<1>::kAtBoundary) d.block_header();
        if (d.state == Inflater<1>::kDone) break;
        if (d.state == Inflater<1>::kAtBoundary) continue;
        uint64_t bp = d.in.pos();
        int32_t pos = d.pos;
        const int rc = warp_symbols_emul(w, size * 8, tab, out, bp, pos, d.floor, (int32_t)out_cap);
        d.pos = pos;
        d.in.seek(w, bp);
        if (rc == kRoundEob) d.end_of_block();
        else if (rc == kRoundBad) d.refuse(kPieceBadData);
        else if (rc == kRoundOverflow) d.refuse(kPieceOverflow);
        else {
            d.symbol();
            while (d.fill_left) d.write_fill();
        }
    }
}

static void inflate_one(bool warp, const uint32_t *w, uint64_t size, uint64_t start_bit, uint64_t stop_bit, int floor0, uint16_t *out,
                        uint32_t out_cap, Tab<1> tab, PieceResult &res)
{
    if (warp) warp_inflate_piece_emul(w, size, start_bit, stop_bit, floor0, out, out_cap, tab, res);
    else inflate_piece(w, size, start_bit, stop_bit, floor0, out, out_cap, tab, res);
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: inflate_emul FILE.gz [piece_bytes [expand [group]]]\n"); return 2; }
    const uint64_t P = argc > 2 ? strtoull(argv[2], nullptr, 10) : 32768;
    const uint64_t expand = argc > 3 ? strtoull(argv[3], nullptr, 10) : 10;
    const size_t G = argc > 4 ? strtoull(argv[4], nullptr, 10) : 64;
    const bool warp = getenv("KIDZ_WARP") != nullptr; // the warp-per-piece decoder instead of the plain one
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    std::vector<uint32_t> words;
    std::vector<uint8_t> raw;
    {
        uint8_t buf[1 << 16];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) raw.insert(raw.end(), buf, buf + n);
        fclose(f);
    }
    const uint64_t size = raw.size();
    words.assign((size + 64 + 3) / 4, 0);
    memcpy(words.data(), raw.data(), size);
    const uint32_t *w = words.data();

    // what zlib says
    std::vector<uint8_t> want;
    bool zlib_ok = true;
    {
        gzFile g = gzopen(argv[1], "rb");
        std::vector<uint8_t> buf(1 << 20);
        int n;
        while ((n = gzread(g, buf.data(), (unsigned)buf.size())) > 0) want.insert(want.end(), buf.begin(), buf.begin() + n);
        if (n < 0) zlib_ok = false;
        if (gzclose(g) != Z_OK) zlib_ok = false;
    }

    const uint64_t hl = gzip_header_len((const uint8_t *)w, size, 0);
    if (!hl) { printf("refused: no gzip header\n"); return 3; }
    const uint64_t first_block_bit = hl * 8;
    const size_t n_pieces = (size + P - 1) / P;
    const uint32_t slot = (uint32_t)(P * expand + (256u << 10));
    std::vector<uint16_t> syms((size_t)n_pieces * slot);
    std::vector<PieceResult> res(n_pieces);
    std::vector<uint16_t> tabmem(kTabEntries), findmem(kFindTabEntries);
    Tab<1> tab{ tabmem.data() }, ftab{ findmem.data() };

    size_t n_cand = 0, n_full = 0, n_nostart = 0;
    uint64_t bits_scanned = 0;
    double t0 = now();
    std::vector<uint64_t> found(n_pieces, ~0ull);
    found[0] = first_block_bit;
    for (size_t k = 1; k < n_pieces; k++) {
        const uint64_t from = k * P * 8, to = piece_end_bit(k, P, size);
        for (uint64_t b = from; b < to; b++) {
            bits_scanned++;
            if (!is_block_start_candidate(w, b)) continue;
            n_cand++;
            if (!is_block_start(w, b, ftab, getenv("KIDZ_TEXT_ONLY") != nullptr)) continue;
            n_full++;
            found[k] = b;
            break;
        }
    }
    double t1 = now();
    size_t n_bad = 0, n_over = 0;
    uint64_t total_syms = 0;
    for (size_t k = 0; k < n_pieces; k++) {
        if (found[k] == ~0ull) {
            res[k] = PieceResult();
            res[k].status = kPieceNoStart;
            n_nostart++;
            continue;
        }
        inflate_one(warp, w, size, found[k], piece_end_bit(k, P, size), k == 0 ? 0 : kWin, syms.data() + k * slot, slot, tab, res[k]);
        if (res[k].status == kPieceOk) resolve_copies(syms.data() + k * slot, res[k].n_out);
        if (res[k].status == kPieceBadData) n_bad++;
        if (res[k].status == kPieceOverflow) n_over++;
        if (res[k].status == kPieceOk) total_syms += res[k].n_out;
    }
    double t2 = now();
    Chain chain;
    const char *why = walk_chain(res, P, size, first_block_bit, [&](size_t k, uint64_t start) {
        inflate_one(warp, w, size, start, piece_end_bit(k, P, size), kWin, syms.data() + k * slot, slot, tab, res[k]);
        if (res[k].status == kPieceOk) resolve_copies(syms.data() + k * slot, res[k].n_out);
        if (getenv("KIDZ_DEBUG")) fprintf(stderr, "redo piece %zu from bit %llu: status %u end %llu n_out %u\n", k, (unsigned long long)start, res[k].status, (unsigned long long)res[k].end_bit, res[k].n_out);
        return true;
    }, 16, chain);
    printf("pieces %zu (%llu bytes each): %zu candidates, %zu headers accepted, %zu without a start, %zu bad, %zu overflow; "
           "%.2f bits scanned per piece byte; find %.3f s, inflate %.3f s (%llu symbols)\n",
           n_pieces, (unsigned long long)P, n_cand, n_full, n_nostart, n_bad, n_over, (double)bits_scanned / (double)(size ? size : 1), t1 - t0,
           t2 - t1, (unsigned long long)total_syms);
    if (why) {
        printf("refused: %s (zlib %s)\n", why, zlib_ok ? "accepts the file" : "refuses it too");
        return 3;
    }
    printf("chain: %zu accepted, %zu inflated again, %zu covered, %zu members, %llu bytes of text\n", chain.pieces.size(), chain.n_redo,
           chain.n_covered, chain.members.size(), (unsigned long long)chain.text_off.back());

    // window maps: groups of G accepted pieces.  pm[j] = what the window before piece j is in terms of the
    // window before its group (a byte, or 256 + index into the group's window); gw[g] = the group's window
    const size_t M = chain.pieces.size(), NG = (M + G - 1) / G;
    std::vector<std::vector<uint16_t>> pm(M), gm(NG);
    uint64_t n_markers = 0;
    for (size_t g = 0; g < NG; g++) {
        std::vector<uint16_t> cur(kWin), nxt(kWin);
        for (int i = 0; i < kWin; i++) cur[i] = (uint16_t)(256 + i);
        for (size_t j = g * G; j < std::min(M, (g + 1) * G); j++) {
            pm[j] = cur;
            const uint32_t k = chain.pieces[j];
            const uint16_t *s = syms.data() + (size_t)k * slot;
            const uint32_t n = res[k].n_out;
            for (int i = 0; i < kWin; i++) {
                if (n < (uint32_t)kWin && (uint32_t)i < kWin - n) { nxt[i] = cur[i + n]; continue; }
                const uint16_t v = s[n - kWin + i];
                nxt[i] = v < 256 ? v : cur[v - 256];
            }
            cur.swap(nxt);
        }
        gm[g] = cur;
    }
    std::vector<std::vector<uint8_t>> gw(NG + 1, std::vector<uint8_t>(kWin, 0));
    for (size_t g = 0; g < NG; g++)
        for (int i = 0; i < kWin; i++) gw[g + 1][i] = gm[g][i] < 256 ? (uint8_t)gm[g][i] : gw[g][gm[g][i] - 256];
    std::vector<uint8_t> text(chain.text_off.back());
    for (size_t j = 0; j < M; j++) {
        const uint32_t k = chain.pieces[j];
        const uint16_t *s = syms.data() + (size_t)k * slot;
        uint8_t *o = text.data() + chain.text_off[j];
        const std::vector<uint8_t> &win = gw[j / G];
        for (uint32_t i = 0; i < res[k].n_out; i++) {
            uint16_t v = s[i];
            if (v >= 256) {
                n_markers++;
                v = pm[j][v - 256];
                if (v >= 256) v = win[v - 256];
            }
            o[i] = (uint8_t)v;
        }
    }
    // CRC-32 per member from 4 KiB chunks, shifted and xor-ed
    uint32_t pow2[32], crctab[256];
    crc_make_pow2(pow2);
    crc_make_table(crctab);
    bool crc_ok = true;
    {
        std::vector<uint32_t> acc(chain.members.size(), 0);
        const uint64_t T = text.size();
        size_t m = 0;
        for (uint64_t c0 = 0; c0 < T; c0 += 4096) {
            uint64_t a = c0;
            const uint64_t cend = std::min(T, c0 + 4096);
            while (a < cend) {
                while (chain.members[m].end <= a) m++;
                const uint64_t b = std::min(cend, chain.members[m].end);
                uint32_t crc = 0xffffffffu;
                for (uint64_t i = a; i < b; i++) crc = crctab[(crc ^ text[i]) & 0xff] ^ (crc >> 8);
                crc ^= 0xffffffffu;
                acc[m] ^= crc_mulmod(crc_x8n(chain.members[m].end - b, pow2), crc);
                a = b;
            }
        }
        for (size_t i = 0; i < acc.size(); i++)
            if (acc[i] != chain.members[i].crc) crc_ok = false;
    }
    const bool same = zlib_ok && text.size() == want.size() && (text.empty() || memcmp(text.data(), want.data(), text.size()) == 0);
    printf("markers %llu (%.3f %% of the text); crc %s; text %s zlib's (%zu bytes)\n", (unsigned long long)n_markers,
           text.empty() ? 0.0 : 100.0 * (double)n_markers / (double)text.size(), crc_ok ? "ok" : "MISMATCH",
           same ? "identical to" : "DIFFERS from", want.size());
    if (!crc_ok) return zlib_ok ? 1 : 3; // a bad check value is a refusal; zlib must refuse such a file too
    return same ? 0 : 1;
}
