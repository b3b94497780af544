// kid_synth - writes the synthetic workload of kid_synth.h in the reference's on-disk formats:
//   kid_synth db    --golden DIR --out WORKDIR [--num 1 --den 100 --seed 10]
//        -> WORKDIR/bact10/{bData10.txt, btree_10.txt, refkey10.txt, probes10.txt.gz}
//   kid_synth reads --golden DIR --out FASTQDIR --sample NAME --pairs N [--first-pair 0 --len 150
//                    --num 1 --den 100 --seed 10 --read-seed 21 --members 65536 --level 1]
//        -> FASTQDIR/NAME_R1_tr.fastq.gz, NAME_R2_tr.fastq.gz   (multi-member gzip, LF, '@S.<pair>/1|2')
// BENCH / TEST TOOLING, not part of the product library.  Probe line format as written by the
// reference's builder (kmer_build_vf6.cpp:625): SEQ,target,org,position,F|R,count.
#include "kid_synth.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h>

static std::map<std::string, std::string> parse_args(int argc, char **argv, int from)
{
    std::map<std::string, std::string> m;
    for (int i = from; i + 1 < argc; i += 2) m[argv[i]] = argv[i + 1];
    return m;
}

static std::string arg(const std::map<std::string, std::string> &m, const char *k, const char *def)
{
    auto it = m.find(k);
    if (it == m.end()) {
        if (!def) { fprintf(stderr, "kid_synth: missing %s\n", k); exit(2); }
        return def;
    }
    return it->second;
}

struct Taxonomy {
    std::vector<int32_t> parent;
    std::vector<uint64_t> prefix;
};

static Taxonomy load_taxonomy(const std::string &golden, uint64_t num, uint64_t den)
{
    Taxonomy t;
    std::vector<uint64_t> counts;
    {
        std::ifstream f(golden + "/refkey10.txt");
        if (!f) { fprintf(stderr, "kid_synth: cannot open %s/refkey10.txt\n", golden.c_str()); exit(2); }
        std::string line;
        std::getline(f, line); // header
        while (std::getline(f, line)) {
            std::vector<std::string> col;
            std::stringstream ss(line);
            std::string c;
            while (std::getline(ss, c, '\t')) col.push_back(c);
            if (col.size() >= 3) counts.push_back(strtoull(col[2].c_str(), nullptr, 10));
        }
    }
    const size_t n = counts.size();
    t.parent.assign(n, 1);
    {
        std::ifstream f(golden + "/btree_10.txt");
        long a, b;
        while (f >> a >> b)
            if (b >= 0 && (size_t)b < n) t.parent[(size_t)b] = (int32_t)a;
    }
    t.prefix.assign(n + 1, 0);
    for (size_t i = 0; i < n; i++) {
        const uint64_t c = (i > 1) ? counts[i] * num / den : 0; // taxa 0 and 1 carry no probes
        t.prefix[i + 1] = t.prefix[i] + c;
    }
    return t;
}

static std::string gzip_member(const std::string &raw, int level)
{
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY);
    std::string out(deflateBound(&zs, raw.size()) + 64, '\0');
    zs.next_in = (Bytef *)raw.data();
    zs.avail_in = (uInt)raw.size();
    zs.next_out = (Bytef *)&out[0];
    zs.avail_out = (uInt)out.size();
    deflate(&zs, Z_FINISH);
    out.resize(zs.total_out);
    deflateEnd(&zs);
    return out;
}

// produce `n_members` gzip members in parallel (make(i) -> raw text) and append them in order
template <class F>
static void write_members(const std::string &path, uint64_t n_members, int level, F make)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "kid_synth: cannot write %s\n", path.c_str()); exit(2); }
    const unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    for (uint64_t base = 0; base < n_members; base += nt) {
        const unsigned cnt = (unsigned)std::min<uint64_t>(nt, n_members - base);
        std::vector<std::string> out(cnt);
        std::vector<std::thread> th;
        for (unsigned k = 0; k < cnt; k++)
            th.emplace_back([&, k] { out[k] = gzip_member(make(base + k), level); });
        for (auto &t : th) t.join();
        for (auto &s : out) fwrite(s.data(), 1, s.size(), f);
    }
    if (n_members == 0) { std::string e = gzip_member("", level); fwrite(e.data(), 1, e.size(), f); }
    fclose(f);
}

static void copy_file(const std::string &a, const std::string &b)
{
    std::ifstream in(a, std::ios::binary);
    std::ofstream out(b, std::ios::binary);
    out << in.rdbuf();
}

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: kid_synth db|reads --key value ...\n"); return 2; }
    const std::string cmd = argv[1];
    auto a = parse_args(argc, argv, 2);
    const std::string golden = arg(a, "--golden", nullptr);
    const uint64_t num = strtoull(arg(a, "--num", "1").c_str(), nullptr, 10);
    const uint64_t den = strtoull(arg(a, "--den", "100").c_str(), nullptr, 10);
    const uint64_t seed = strtoull(arg(a, "--seed", "10").c_str(), nullptr, 10);
    const int level = atoi(arg(a, "--level", "1").c_str());
    Taxonomy tax = load_taxonomy(golden, num, den);
    const uint32_t n_taxa = (uint32_t)tax.parent.size();
    const uint64_t n_probes = tax.prefix.back();

    if (cmd == "db") {
        const std::string out = arg(a, "--out", nullptr) + "/bact10";
        std::string mk = "mkdir -p '" + out + "'";
        if (system(mk.c_str()) != 0) return 2;
        copy_file(golden + "/bData10.txt", out + "/bData10.txt");
        copy_file(golden + "/btree_10.txt", out + "/btree_10.txt");
        copy_file(golden + "/refkey10.txt", out + "/refkey10.txt");
        const uint64_t per = 1 << 18;
        write_members(out + "/probes10.txt.gz", (n_probes + per - 1) / per, level, [&](uint64_t mi) {
            std::string s;
            s.reserve(per * 52);
            char buf[96];
            for (uint64_t i = mi * per; i < std::min(n_probes, (mi + 1) * per); i++) {
                const uint64_t k = ks_probe_key(seed, i);
                for (int b = 29; b >= 0; b--) buf[29 - b] = "ACGT"[(k >> (2 * b)) & 3];
                const uint32_t t = ks_taxon_of(tax.prefix.data(), n_taxa, i);
                const int len = snprintf(buf + 30, sizeof buf - 30, ",%u,%u,%u,%c,1\n", t,
                                         (unsigned)(i % 14791), (unsigned)(i % 100000), (i & 1) ? 'F' : 'R');
                s.append(buf, 30 + (size_t)len);
            }
            return s;
        });
        printf("%llu probes -> %s/probes10.txt.gz\n", (unsigned long long)n_probes, out.c_str());
        return 0;
    }
    if (cmd == "reads") {
        const std::string out = arg(a, "--out", nullptr);
        const std::string sample = arg(a, "--sample", nullptr);
        const uint64_t pairs = strtoull(arg(a, "--pairs", nullptr).c_str(), nullptr, 10);
        const uint64_t first = strtoull(arg(a, "--first-pair", "0").c_str(), nullptr, 10);
        const uint64_t per = strtoull(arg(a, "--members", "65536").c_str(), nullptr, 10);
        ks_config c;
        memset(&c, 0, sizeof c);
        c.seed_db = seed;
        c.seed_reads = strtoull(arg(a, "--read-seed", "21").c_str(), nullptr, 10);
        c.n_probes = n_probes;
        c.n_taxa = n_taxa;
        c.read_len = (uint32_t)atoi(arg(a, "--len", "150").c_str());
        c.stride = c.read_len;
        c.on_target_pct = (uint32_t)atoi(arg(a, "--on-target", "70").c_str());
        c.sub_per_10k = 50;
        c.n_per_10k = 10;
        c.bad_tail_pct = 20;
        std::string mk = "mkdir -p '" + out + "'";
        if (system(mk.c_str()) != 0) return 2;
        for (int mate = 0; mate < 2; mate++) {
            const std::string path = out + "/" + sample + (mate ? "_R2_tr.fastq.gz" : "_R1_tr.fastq.gz");
            write_members(path, (pairs + per - 1) / per, level, [&](uint64_t mi) {
                std::string s;
                std::vector<uint8_t> sq(c.read_len), ql(c.read_len);
                char name[64];
                for (uint64_t p = mi * per; p < std::min(pairs, (mi + 1) * per); p++) {
                    const uint64_t g = 2 * (first + p) + (uint64_t)mate;
                    ks_make_read(&c, tax.prefix.data(), tax.parent.data(), g, sq.data(), ql.data());
                    const int nl = snprintf(name, sizeof name, "@S.%llu/%d\n", (unsigned long long)(first + p), mate + 1);
                    s.append(name, (size_t)nl);
                    s.append((const char *)sq.data(), c.read_len);
                    s.append("\n+\n");
                    s.append((const char *)ql.data(), c.read_len);
                    s.push_back('\n');
                }
                return s;
            });
        }
        printf("%llu pairs -> %s/%s_R[12]_tr.fastq.gz\n", (unsigned long long)pairs, out.c_str(), sample.c_str());
        return 0;
    }
    fprintf(stderr, "kid_synth: unknown command %s\n", cmd.c_str());
    return 2;
}
