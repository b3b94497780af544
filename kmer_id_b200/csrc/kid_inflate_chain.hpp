// kid_inflate_chain.hpp - the host's part of the device-side inflate (kid_inflate.cuh): pieces are
// accepted in file order only if each started exactly where its predecessor stopped, so that, by
// induction from the first bit of the file, every accepted start is a real block boundary (the rule of
// host/pgz.cpp).  A piece that started anywhere else (or nowhere) is inflated again from its
// predecessor's end through `redo`; pieces a predecessor ran past are dropped.
#pragma once
#include "kid_inflate.cuh"

#include <algorithm>
#include <vector>

namespace kidz {

struct Member { // one gzip member in the inflated text
    uint64_t begin, end;
    uint32_t crc, isize;
};

struct Chain {
    std::vector<uint32_t> pieces;   // accepted pieces, stream order
    std::vector<uint64_t> text_off; // pieces.size() + 1: where each piece's bytes go
    std::vector<Member> members;
    size_t n_redo = 0, n_covered = 0;
};

inline uint64_t piece_end_bit(size_t k, uint64_t piece_bytes, uint64_t size) { return std::min<uint64_t>(size, (k + 1) * piece_bytes) * 8; }

// redo(k, start_bit) inflates piece k again from start_bit and refreshes res[k]; false = cannot.
// Returns nullptr on success, else why the file is not one for this decoder.
template <class Redo>
const char *walk_chain(std::vector<PieceResult> &res, uint64_t piece_bytes, uint64_t size, uint64_t first_block_bit, Redo redo,
                       size_t max_redo, Chain &c)
{
    c = Chain();
    uint64_t cur = first_block_bit, text = 0, member_begin = 0;
    bool eof = false;
    c.text_off.push_back(0);
    for (size_t k = 0; k < res.size(); k++) {
        if (eof || (k > 0 && cur >= piece_end_bit(k, piece_bytes, size))) { // a predecessor covered it
            c.n_covered++;
            continue;
        }
        if (!(res[k].status == kPieceOk && res[k].start_bit == cur)) {
            if (c.n_redo >= max_redo) return "too many pieces to inflate again";
            c.n_redo++;
            if (!redo(k, cur) || res[k].status != kPieceOk || res[k].start_bit != cur) return "a piece does not inflate";
        }
        const PieceResult &r = res[k];
        for (uint32_t e = 0; e < r.n_ends; e++) {
            c.members.push_back(Member{ member_begin, text + r.ends[e].out_pos, r.ends[e].crc, r.ends[e].isize });
            member_begin = text + r.ends[e].out_pos;
        }
        c.pieces.push_back((uint32_t)k);
        text += r.n_out;
        c.text_off.push_back(text);
        cur = r.end_bit;
        eof = r.eof != 0;
    }
    if (!eof) return "the stream does not end with a complete member";
    if (member_begin != text) return "data after the last member";
    for (const Member &m : c.members)
        if ((uint32_t)(m.end - m.begin) != m.isize) return "a member's length does not match its trailer";
    return nullptr;
}

} // namespace kidz
