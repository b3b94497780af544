// kmerreadc - GPU drop-in for the reference's fungal/viral reader (kmer_read_vf6.cpp main():965-1172).
//
// Same contract:  kmerreadc -name <db> -jname <jobs> [-target <taxon>] [-fadir <dir>]
//   ./<db>/<db>_data.txt, <db>_tree.txt, <db>_probes.txt.gz ; job list ./<jobs>/<jobs>.txt
//   ("<job> <n>" followed by n file names, :1022-1053); per job ./<jobs>/<job>_result.txt,
//   <job>_reads.txt (first 12 reads per taxon, only when -target is absent/0, :613-616) and
//   <job>_target_reads.txt (every read of taxon -target, :617-620); same stdout lines.
// Reads accept U/u as T (:496-500,521-525 -> KID_DB_ACCEPT_U).  -fadir only feeds the dead
// Smith-Waterman branch and is ignored.
#include "../../include/kmer_id.h"
#include "db_loader.hpp"
#include "device_warmup.hpp"
#include "batch_pipeline.hpp"
#include "read_reader.hpp"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

using namespace kidhost;

namespace {
const int SAVENUM = 12; // kmer_read_vf6.cpp:40

[[noreturn]] void die(int code, const std::string &msg)
{
    std::cerr << "kmerreadc: " << msg << std::endl;
    exit(code);
}

bool ends_with(const std::string &s, const std::string &suffix)
{
    return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}

struct JobState {
    std::vector<int> gcount_host;
    long long tct = 0;
    int save_target = 0;
    std::ofstream outread1, outread2;
};

void run_file(kid_sample *smp, ReadFormat fmt, const std::string &path, JobState &st)
{
    ReadBatchReader reader(fmt, path, (size_t)1 << 19, (size_t)96 << 20, pipeline_batches(), 0, BatchMode::Packed, KID_DB_ACCEPT_U);
    classify_stream(smp, reader, [&](const ReadBatch &b) {
        for (size_t r = 0; r < b.n; r++) { // process_read :613-623, in stream order
            const int fin = b.taxon[r];
            if (fin < 0) continue;
            const bool first12 = fin > 1 && st.gcount_host[(size_t)fin] < SAVENUM && st.save_target == 0;
            const bool wanted = fin > 1 && fin == st.save_target;
            for (int k = 0; k < 2; k++) {
                if (!(k == 0 ? first12 : wanted)) continue;
                std::ofstream &o = k == 0 ? st.outread1 : st.outread2;
                o << ">" << fin << ":";
                o.write(b.names.data() + b.name_off[r], b.name_off[r + 1] - b.name_off[r]);
                o << std::endl;
                o.write((const char *)b.seq + b.off[r] + b.span[2 * r], b.span[2 * r + 1] - b.span[2 * r] + 1);
                o << std::endl;
            }
            st.gcount_host[(size_t)fin]++;
            st.tct++;
        }
    });
    if (fmt == ReadFormat::PlainFasta && reader.open_failed()) std::cout << "nark " << path << std::endl;
}
} // namespace

int main(int argc, char *argv[])
{
    std::string dname, wdir, jname, jdir;
    int save_target = 0;
    for (int i = 1; i + 1 < argc; i++) { // :986-1014
        const std::string a = argv[i];
        if (a == "-name") { dname = argv[i + 1]; wdir = "./" + dname + "/"; }
        if (a == "-jname") { jname = argv[i + 1]; jdir = "./" + jname + "/"; }
        if (a == "-target") save_target = atoi(argv[i + 1]);
    }
    const int device = getenv("KID_DEVICE") ? atoi(getenv("KID_DEVICE")) : 0;
    const std::string iname = wdir + dname + "_data.txt", tname = wdir + dname + "_tree.txt",
                      pname = wdir + dname + "_probes.txt.gz", jfile = jdir + jname + ".txt";

    // job list (:1022-1053): "<job> <n>" followed by n lines whose first token is a file name
    std::vector<std::string> jnames;
    std::vector<std::vector<std::string>> fnames;
    {
        std::ifstream fin(jfile);
        if (fin) {
            std::string line;
            while (std::getline(fin, line)) {
                if (!line.empty() && line.back() == '\r') line.pop_back();
                if (line.length() <= 1) continue;
                std::stringstream ls(line);
                std::string name, tok;
                int j = 0;
                ls >> name >> j;
                // for j <= 0 the reference's jnames/fnames indices drift apart and the next job reads
                // fnames[][] out of range (:1033-1047): refuse instead of guessing
                if (j <= 0) die(1, "job list: '" + name + "' has no files (undefined behaviour in the reference)");
                std::vector<std::string> files;
                tok = name;
                for (int i = 0; i < j; i++) {
                    std::getline(fin, line);
                    if (!line.empty() && line.back() == '\r') line.pop_back();
                    std::stringstream fs(line);
                    fs >> tok; // a failed extraction keeps the previous token, like the reference's jstr
                    files.push_back(tok);
                }
                jnames.push_back(name);
                fnames.push_back(files);
            }
            std::cout << jnames.size() << " jobs" << std::endl;
        } else {
            std::cout << "narin " << jfile << std::endl;
        }
    }

    // strain list (:1055-1085)
    int num_targ = 0, num_orgs = 0;
    {
        std::ifstream fin(iname);
        if (fin) {
            std::string line, acc;
            int targi = 0;
            while (std::getline(fin, line)) {
                if (!line.empty() && line.back() == '\r') line.pop_back();
                if (line.length() > 1) {
                    std::stringstream ls(line);
                    ls >> targi >> acc;
                    if (targi > num_targ) num_targ = targi;
                    num_orgs++;
                }
            }
            std::cout << num_orgs << " strains" << std::endl;
            std::cout << num_targ << " targs" << std::endl;
            num_targ++;
        } else {
            std::cout << "narin " << iname << std::endl;
        }
    }
    if (num_targ < 2) die(1, "no usable " + iname + " (the reference would build an empty taxonomy)");

    std::vector<int32_t> parent((size_t)num_targ, 1);
    {
        std::ifstream fin(tname); // a missing tree is silently accepted (:1089-1102)
        std::string line;
        int i = 0, j = 0;
        while (fin && std::getline(fin, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            std::stringstream ls(line);
            ls >> i >> j;
            if (i < 0 || i >= num_targ || j < 0 || j >= num_targ)
                die(1, "taxonomy edge " + std::to_string(i) + " " + std::to_string(j) + " is outside [0," +
                           std::to_string(num_targ) + ")");
            parent[(size_t)j] = i;
        }
    }
    std::cout << "tree loaded" << std::endl;

    ProbeSet probes;
    start_device_warmup(device); // CUDA context creation overlaps the parse
    load_probes_cached(pname, probes, /*target_signed=*/false);
    finish_device_warmup();
    std::cout << probes.lines_parsed << " kmers loaded" << std::endl;
    kid_db *db = nullptr;
    if (kid_db_build(probes.keys.data(), probes.taxa.data(), probes.keys.size(), 0, parent.data(), num_targ, device,
                     KID_DB_ACCEPT_U, 0, nullptr, &db) != 0)
        die(1, kid_last_error());
    { std::vector<uint64_t>().swap(probes.keys); std::vector<uint32_t>().swap(probes.taxa); }
    kid_sample *smp = nullptr;
    if (kid_sample_create(db, &smp) != 0) die(1, kid_last_error());

    std::vector<int32_t> gcount((size_t)num_targ), ucount((size_t)num_targ);
    for (size_t jb = 0; jb < jnames.size(); jb++) { // :1110-1160
        if (kid_sample_begin(smp, nullptr) != 0) die(1, kid_last_error());
        JobState st;
        st.gcount_host.assign((size_t)num_targ, 0);
        st.save_target = save_target;
        const std::string base = "./" + jname + "/" + jnames[jb];
        st.outread1.open((base + "_reads.txt").c_str(), std::ofstream::out | std::ofstream::trunc);
        if (save_target > 0) st.outread2.open((base + "_target_reads.txt").c_str(), std::ofstream::out | std::ofstream::trunc);
        for (const std::string &r1name : fnames[jb]) {
            std::cout << r1name << std::endl;
            if (ends_with(r1name, ".fastq.gz")) run_file(smp, ReadFormat::GzFastq, r1name, st);
            else if (ends_with(r1name, ".fasta.gz")) run_file(smp, ReadFormat::GzFasta, r1name, st);
            else if (ends_with(r1name, ".fasta")) run_file(smp, ReadFormat::PlainFasta, r1name, st);
            else if (ends_with(r1name, ".fastq")) run_file(smp, ReadFormat::PlainFastq, r1name, st);
        }
        std::cout << st.tct << " reads loaded" << std::endl;
        st.outread1.close();
        if (save_target > 0) st.outread2.close();
        if (kid_sample_counts(smp, gcount.data(), ucount.data(), nullptr) != 0) die(1, kid_last_error());
        std::ofstream out2(base + "_result.txt");
        for (int i = 0; i < num_targ; i++) out2 << i << "," << gcount[(size_t)i] << "," << ucount[(size_t)i] << "\n";
        out2.close();
    }
    kid_sample_free(smp);
    kid_db_free(db);
    return 0;
}
