#include "db_loader.hpp"

#include <new>
#include "gz_lines.hpp"

#include <algorithm>
#include <condition_variable>
#include <deque>
#include <fstream>
#include <memory>
#include <mutex>
#include <sstream>
#include <thread>

#include <sys/stat.h>
#include <unistd.h>

namespace kidhost {

bool load_tree(const std::string &path, int n_taxa, std::vector<int32_t> &parent, std::string &msg)
{
    parent.assign((size_t)n_taxa, 1); // Tree1(): every node hangs off the root
    std::ifstream fin(path);
    if (!fin) return true; // the reference prints "tree loaded" regardless (:974-984)
    std::string line;
    int i = 0, j = 0; // carried across lines when an extraction fails, like the reference's locals
    while (std::getline(fin, line)) {
        std::stringstream ls(line);
        ls >> i >> j;
        if (i < 0 || i >= n_taxa || j < 0 || j >= n_taxa) {
            msg = "taxonomy edge " + std::to_string(i) + " " + std::to_string(j) +
                  " is outside [0," + std::to_string(n_taxa) + ")";
            return false;
        }
        parent[(size_t)j] = i;
    }
    return true;
}

static inline bool is_ws(unsigned char c)
{
    return c == ' ' || (c >= '\t' && c <= '\r');
}

// process_kmer (:619-661): forward strand, upper-case ACGT only, every full 30-window
static const struct BaseCodes {
    uint8_t code[256];
    BaseCodes()
    {
        for (int i = 0; i < 256; i++) code[i] = 0xFF;
        code['A'] = 0; code['C'] = 1; code['G'] = 2; code['T'] = 3;
    }
} kBase;

static inline void add_windows(const char *s, size_t n, uint32_t target, std::vector<uint64_t> &keys,
                               std::vector<uint32_t> &taxa)
{
    const uint64_t mask = (1ULL << 60) - 1;
    uint64_t kf = 0;
    int cpos = 0;
    for (size_t i = 0; i < n; i++) {
        const unsigned c = kBase.code[(unsigned char)s[i]]; // table, not a switch: the bases are random
        if (c > 3) { cpos = 0; kf = 0; continue; }
        kf = ((kf << 2) & mask) | c;
        if (++cpos == 30) {
            keys.push_back(kf);
            taxa.push_back(target);
            cpos--;
        }
    }
}

// digits only, 1..9 of them (cannot overflow an int); anything else sends the line to the slow path
static inline bool fast_uint(const char *&p, const char *end, uint32_t &v)
{
    const char *q = p;
    uint32_t x = 0;
    while (q < end && *q >= '0' && *q <= '9' && q - p < 10) x = x * 10 + (uint32_t)(*q++ - '0');
    if (q == p || q - p > 9) return false;
    v = x;
    p = q;
    return true;
}

bool parse_probe_line(const char *line, size_t len, std::vector<uint64_t> &keys,
                      std::vector<uint32_t> &taxa, bool target_signed)
{
    if (len > 0 && line[len - 1] == '\r') len--; // :691-692
    if (len == 0) return false;                   // :693
    const char *p = line, *end = line + len;
    // fast path: SEQ,uint,uint,uint,char,uint  - the format the builder writes.  One pass over SEQ
    // makes the window keys (process_kmer :619-661); they are kept only if the rest of the line parses.
    auto rest_ok = [&](const char *q, uint32_t &target) { // ",uint,uint,uint,char,uint" from the first comma
        uint32_t org, pos, count;
        return q < end && *q == ',' && (++q, fast_uint(q, end, target)) && q < end && *q == ',' &&
               (++q, fast_uint(q, end, org)) && q < end && *q == ',' && (++q, fast_uint(q, end, pos)) &&
               q < end && *q == ',' && q + 2 < end && q[1] != ',' && !is_ws((unsigned char)q[1]) &&
               q[2] == ',' && (q += 3, fast_uint(q, end, count));
    };
    if (len > 31 && line[30] == ',') { // the usual line: exactly one 30-mer, two independent halves
        uint64_t a = 0, b = 0;
        unsigned bad = 0;
        for (int i = 0; i < 15; i++) {
            const unsigned ca = kBase.code[(unsigned char)line[i]], cb = kBase.code[(unsigned char)line[15 + i]];
            bad |= ca | cb;
            a = (a << 2) | ca;
            b = (b << 2) | cb;
        }
        uint32_t target;
        if (bad <= 3 && rest_ok(line + 30, target)) {
            keys.push_back((a << 30) | b);
            taxa.push_back(target);
            return true;
        }
    }
    {
        const size_t k0 = keys.size();
        const uint64_t mask = (1ULL << 60) - 1;
        uint64_t kf = 0;
        int cpos = 0;
        for (; p < end; p++) {
            const unsigned c = kBase.code[(unsigned char)*p];
            if (c <= 3) {
                kf = ((kf << 2) & mask) | c;
                if (++cpos == 30) { keys.push_back(kf); cpos--; }
                continue;
            }
            if (*p == ',' || is_ws((unsigned char)*p)) break;
            cpos = 0; // any other character restarts the window
            kf = 0;
        }
        uint32_t target;
        if (p > line && rest_ok(p, target)) {
            taxa.insert(taxa.end(), keys.size() - k0, target);
            return true;
        }
        keys.resize(k0);
    }
    // slow path: the reference's own extraction sequence on the comma->blank line (:695-697)
    std::string l(line, len);
    std::replace(l.begin(), l.end(), ',', ' ');
    std::istringstream ss(l);
    std::string sequence;
    int org, position, count;
    char strand;
    if (target_signed) {
        int target;
        if (ss >> sequence >> target >> org >> position >> strand >> count) {
            add_windows(sequence.data(), sequence.size(), (uint32_t)target, keys, taxa);
            return true;
        }
    } else {
        unsigned int target;
        if (ss >> sequence >> target >> org >> position >> strand >> count) {
            add_windows(sequence.data(), sequence.size(), target, keys, taxa);
            return true;
        }
    }
    return false;
}

namespace {
struct Block {
    LineBlock text;
    std::vector<uint64_t> keys;
    std::vector<uint32_t> taxa;
    long long lines = 0;
    bool done = false;
};
} // namespace

void load_probes_gz(const std::string &path, ProbeSet &out, bool target_signed, unsigned threads)
{
    if (threads == 0) threads = std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
    GzLineBlocks src(path);
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::shared_ptr<Block>> todo, inorder;
    bool eof = false;

    auto worker = [&] {
        for (;;) {
            std::shared_ptr<Block> b;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !todo.empty() || eof; });
                if (todo.empty()) return;
                b = todo.front();
                todo.pop_front();
            }
            b->keys.reserve(b->text.size() / 48);
            b->taxa.reserve(b->text.size() / 48);
            auto parse_lines = [&](const char *p, const char *end) {
                while (p < end) {
                    const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
                    if (!eol) break;
                    b->lines += parse_probe_line(p, (size_t)(eol - p), b->keys, b->taxa, target_signed);
                    p = eol + 1;
                }
            };
            parse_lines(b->text.head.data(), b->text.head.data() + b->text.head.size());
            parse_lines(b->text.body, b->text.body + b->text.body_len);
            b->text = LineBlock(); // hands the inflated buffer back
            {
                std::lock_guard<std::mutex> lk(mu);
                b->done = true;
            }
            cv.notify_all();
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; t++) pool.emplace_back(worker);

    auto drain = [&](bool all) { // append finished blocks in file order
        for (;;) {
            std::shared_ptr<Block> b;
            {
                std::unique_lock<std::mutex> lk(mu);
                if (inorder.empty()) return;
                if (all) cv.wait(lk, [&] { return inorder.front()->done; });
                if (!inorder.front()->done) return;
                b = inorder.front();
                inorder.pop_front();
            }
            out.keys.insert(out.keys.end(), b->keys.begin(), b->keys.end());
            out.taxa.insert(out.taxa.end(), b->taxa.begin(), b->taxa.end());
            out.lines_parsed += b->lines;
        }
    };

    // One key per ~50 bytes of text in the usual file; reserving address space up front (untouched
    // pages cost nothing) saves the copies of a growing vector.  More keys than that just grow it.
    {
        struct stat st;
        if (stat(path.c_str(), &st) == 0 && st.st_size > 0) {
            const size_t guess = (size_t)st.st_size / 10 + 1024;
            out.keys.reserve(out.keys.size() + guess);
            out.taxa.reserve(out.taxa.size() + guess);
        }
    }
    for (;;) {
        auto b = std::make_shared<Block>();
        if (!src.next(b->text)) break;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return todo.size() < 4 * threads; });
            todo.push_back(b);
            inorder.push_back(b);
        }
        cv.notify_all();
        drain(false);
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        eof = true;
    }
    cv.notify_all();
    drain(true);
    for (auto &t : pool) t.join();
}


namespace {
struct CacheHeader {
    char magic[8];       // "KIDCACH1"
    uint64_t src_size;   // st_size of the text file
    int64_t src_mtime_ns;
    uint64_t n_entries;
    int64_t lines_parsed;
    uint32_t target_signed;
    uint32_t reserved;
};
} // namespace

bool load_probes_cached(const std::string &path, ProbeSet &out, bool target_signed)
{
    struct stat st;
    const bool have_src = stat(path.c_str(), &st) == 0;
    const std::string cpath = path + ".kidcache";
    const bool allow = getenv("KID_NO_CACHE") == nullptr;
    if (have_src && allow) {
        if (FILE *f = fopen(cpath.c_str(), "rb")) {
            CacheHeader h;
            bool ok = fread(&h, sizeof h, 1, f) == 1 && memcmp(h.magic, "KIDCACH1", 8) == 0 &&
                      h.src_size == (uint64_t)st.st_size &&
                      h.src_mtime_ns == (int64_t)st.st_mtim.tv_sec * 1000000000LL + st.st_mtim.tv_nsec &&
                      h.target_signed == (uint32_t)target_signed;
            // a truncated or damaged cache must fall back to the text, not die in resize()
            struct stat cst;
            ok = ok && fstat(fileno(f), &cst) == 0 && h.n_entries < 0xFFFFFFFFull &&
                 (uint64_t)cst.st_size == sizeof h + 12ull * h.n_entries;
            if (ok) {
                try {
                    out.keys.resize((size_t)h.n_entries);
                    out.taxa.resize((size_t)h.n_entries);
                } catch (const std::bad_alloc &) {
                    ok = false;
                }
            }
            if (ok) {
                ok = fread(out.keys.data(), 8, out.keys.size(), f) == out.keys.size() &&
                     fread(out.taxa.data(), 4, out.taxa.size(), f) == out.taxa.size();
                out.lines_parsed = h.lines_parsed;
            }
            fclose(f);
            if (ok) return true;
            out.keys.clear();
            out.taxa.clear();
            out.lines_parsed = 0;
        }
    }
    load_probes_gz(path, out, target_signed);
    if (have_src && allow) {
        const std::string tmp = cpath + ".tmp" + std::to_string((long)getpid());
        if (FILE *f = fopen(tmp.c_str(), "wb")) {
            CacheHeader h;
            memset(&h, 0, sizeof h);
            memcpy(h.magic, "KIDCACH1", 8);
            h.src_size = (uint64_t)st.st_size;
            h.src_mtime_ns = (int64_t)st.st_mtim.tv_sec * 1000000000LL + st.st_mtim.tv_nsec;
            h.n_entries = out.keys.size();
            h.lines_parsed = out.lines_parsed;
            h.target_signed = (uint32_t)target_signed;
            const bool ok = fwrite(&h, sizeof h, 1, f) == 1 &&
                            fwrite(out.keys.data(), 8, out.keys.size(), f) == out.keys.size() &&
                            fwrite(out.taxa.data(), 4, out.taxa.size(), f) == out.taxa.size();
            if (fclose(f) != 0 || !ok || rename(tmp.c_str(), cpath.c_str()) != 0) remove(tmp.c_str());
        }
    }
    return false;
}

namespace {
// MurmurHash3's 64-bit finaliser: the hash the reference's table uses (kmer_read_m3.cpp:191-199)
inline uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

// insertion depth (number of probes) each entry would need in the reference's table, in file order
template <class F>
void replay_reference_inserts(const ProbeSet &probes, int log2_cells, F on_depth)
{
    const uint64_t mask = (1ULL << log2_cells) - 1;
    std::vector<uint64_t> occupied((size_t)((mask + 1) >> 6), 0);
    for (size_t e = 0; e < probes.keys.size(); e++) {
        if (probes.taxa[e] == 0) continue; // value 0 leaves the cell looking empty (:248-255)
        const uint64_t hash = fmix64(probes.keys[e]);
        uint64_t reprobe = 0, i = 0;
        for (;;) {
            const uint64_t index = (hash + reprobe) & mask;
            reprobe += ++i;
            uint64_t &w = occupied[(size_t)(index >> 6)];
            const uint64_t bit = 1ULL << (index & 63);
            if (!(w & bit)) { w |= bit; break; }
        }
        on_depth(e, (int)i);
    }
}
} // namespace

size_t apply_reference_probe_cap(ProbeSet &probes, int cap, int log2_cells)
{
    std::vector<uint64_t> deep; // keys with at least one copy beyond the cap (practically never any)
    replay_reference_inserts(probes, log2_cells, [&](size_t e, int depth) {
        if (depth > cap) deep.push_back(probes.keys[e]);
    });
    if (deep.empty()) return 0;
    std::sort(deep.begin(), deep.end());
    deep.erase(std::unique(deep.begin(), deep.end()), deep.end());
    auto is_deep = [&](uint64_t k) { return std::binary_search(deep.begin(), deep.end(), k); };
    // depth of the FIRST copy decides: it is the one getHash meets first
    std::vector<int> first_depth(deep.size(), 0);
    replay_reference_inserts(probes, log2_cells, [&](size_t e, int depth) {
        const uint64_t k = probes.keys[e];
        if (!is_deep(k)) return;
        const size_t j = (size_t)(std::lower_bound(deep.begin(), deep.end(), k) - deep.begin());
        if (first_depth[j] == 0) first_depth[j] = depth;
    });
    size_t hidden = 0;
    for (size_t j = 0; j < deep.size(); j++) hidden += first_depth[j] > cap;
    for (size_t e = 0; e < probes.keys.size(); e++) {
        const uint64_t k = probes.keys[e];
        if (!is_deep(k)) continue;
        const size_t j = (size_t)(std::lower_bound(deep.begin(), deep.end(), k) - deep.begin());
        if (first_depth[j] > cap) probes.taxa[e] = 0;
    }
    return hidden;
}

} // namespace kidhost
