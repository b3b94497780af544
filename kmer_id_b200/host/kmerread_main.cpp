// kmerread - GPU drop-in for the reference's mitokmer reader (kmer_read_m3.cpp main():973-1132).
//
// Same contract:  kmerread -wdir <dir/> -f1 <reads> [-f2 <reads|none>]
//   <wdir>mitochondria_data.txt, mitochondria_tree.txt, mitochondria_probes.txt.gz  ->  <wdir>result.txt
//   input kind by suffix: .fastq.gz / .fasta / .fastq / .fasta.gz (:1076-1120); same stdout lines.
// Differences from nk10 that are reproduced: the number of taxa is max(target)+1 of the data file
// (:1036-1044; the reference forgets to initialise num_targ, :981 - we start it at 0 like
// kmer_read_vf6.cpp:974 does), lookups give up after 16 probes (:42,232, see
// apply_reference_probe_cap), no _reads.txt (commented out, :612-621).
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
[[noreturn]] void die(int code, const std::string &msg)
{
    std::cerr << "kmerread: " << msg << std::endl;
    exit(code);
}

bool ends_with(const std::string &s, const std::string &suffix)
{ // equal(suffix.rbegin(), suffix.rend(), s.rbegin()) of the reference, without reading before s
    return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}

long long g_tct = 0;

// gz FASTQ on the device (kid_fastq_*, include/kmer_id.h): inflate, framing, trim and scan in kernels; only tct
// (:624) is kept per read here.  false = a file for the host reader (nothing was counted).  KID_GPU_INGEST=0: never.
bool run_gz_fastq_on_device(kid_sample *smp, const std::string &path)
{
    static const bool enabled = !(getenv("KID_GPU_INGEST") && atoi(getenv("KID_GPU_INGEST")) == 0);
    if (!enabled) return false;
    static kid_fastq *fq = nullptr; // lives until exit, like the sample
    if (!fq && kid_fastq_create(kid_sample_db(smp), &fq) != 0) return false;
    size_t n = 0;
    const int rc = kid_fastq_load_gz_file(fq, path.c_str(), &n);
    if (rc == KID_EUNSUPPORTED) return false;
    if (rc != 0) { std::cerr << "kmerread: " << kid_last_error() << std::endl; exit(1); }
    void *p = nullptr;
    if (kid_host_alloc(&p, sizeof(int32_t) * (n + 1)) != 0) { std::cerr << "kmerread: " << kid_last_error() << std::endl; exit(1); }
    int32_t *taxon = (int32_t *)p;
    if (kid_fastq_classify(fq, smp, taxon) != 0) { std::cerr << "kmerread: " << kid_last_error() << std::endl; exit(1); }
    for (size_t r = 0; r < n; r++) g_tct += taxon[r] >= 0; // tct++ per processed read (:624)
    kid_host_free(p);
    return true;
}

void run_file(kid_sample *smp, ReadFormat fmt, const std::string &path)
{
    if (fmt == ReadFormat::GzFastq && run_gz_fastq_on_device(smp, path)) return;
    if (fmt == ReadFormat::GzFasta) std::cout << "true" << std::endl; // process_fagz :789
    ReadBatchReader reader(fmt, path, (size_t)1 << 19, (size_t)96 << 20, pipeline_batches());
    classify_stream(smp, reader, [&](const ReadBatch &b) {
        for (size_t r = 0; r < b.n; r++) g_tct += b.taxon[r] >= 0; // tct++ per processed read (:624)
    });
    if (fmt == ReadFormat::PlainFasta && reader.open_failed()) std::cout << "nark " << path << std::endl; // :969-971
}

// the suffix dispatch of :1076-1120; returns false if no suffix matched (the reference then does nothing)
bool dispatch(kid_sample *smp, const std::string &name, bool second)
{
    if (ends_with(name, ".fastq.gz")) {
        run_file(smp, ReadFormat::GzFastq, name);
        if (second) std::cout << g_tct << " reads loaded" << std::endl; // :1106
    } else if (ends_with(name, ".fasta")) run_file(smp, ReadFormat::PlainFasta, name);
    else if (ends_with(name, ".fastq")) run_file(smp, ReadFormat::PlainFastq, name);
    else if (ends_with(name, ".fasta.gz")) run_file(smp, ReadFormat::GzFasta, name);
    else return false;
    return true;
}
} // namespace

int main(int argc, char *argv[])
{
    std::string wdir, r1name, r2name;
    for (int i = 1; i + 1 < argc; i++) { // :993-1010
        const std::string a = argv[i];
        if (a == "-wdir") wdir = argv[i + 1];
        if (a == "-f1") r1name = argv[i + 1];
        if (a == "-f2") r2name = argv[i + 1];
    }
    const int device = getenv("KID_DEVICE") ? atoi(getenv("KID_DEVICE")) : 0;
    const std::string iname = wdir + "mitochondria_data.txt", tname = wdir + "mitochondria_tree.txt",
                      pname = wdir + "mitochondria_probes.txt.gz";
    std::cout << "r1 " << r1name << std::endl;
    std::cout << "r2 " << r2name << std::endl;
    std::cout << "wd " << wdir << std::endl;
    if (r1name.empty()) die(134, "-f1 is required (the reference aborts in r1name.at(-1), :1074)");

    // strain list: only max(target) matters on this path (:1024-1044)
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
            num_targ++;
        }
    }
    if (num_targ < 2) die(1, "no usable " + iname + " (the reference would build an empty taxonomy)");

    // taxonomy (:1047-1062); a missing tree file is fatal here, unlike nk10
    std::vector<int32_t> parent((size_t)num_targ, 1);
    {
        std::ifstream fin(tname);
        if (!fin) exit(1);
        std::string line;
        int i = 0, j = 0;
        while (std::getline(fin, line)) {
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
    load_probes_cached(pname, probes, /*target_signed=*/true);
    finish_device_warmup();
    std::cout << probes.lines_parsed << " kmers loaded" << std::endl; // :1066
    if (probes.lines_parsed < 2) exit(1);                              // :1067
    // KID_REF_LOG2_CELLS: test hook - size of the reference table being replayed (MAXHASH, :41)
    apply_reference_probe_cap(probes, 16, getenv("KID_REF_LOG2_CELLS") ? atoi(getenv("KID_REF_LOG2_CELLS")) : 30);
    kid_db *db = nullptr;
    if (kid_db_build(probes.keys.data(), probes.taxa.data(), probes.keys.size(), 0, parent.data(), num_targ, device,
                     0, 0, nullptr, &db) != 0)
        die(1, kid_last_error());
    { std::vector<uint64_t>().swap(probes.keys); std::vector<uint32_t>().swap(probes.taxa); }
    kid_sample *smp = nullptr;
    if (kid_sample_create(db, &smp) != 0) die(1, kid_last_error());

    std::cout << r1name.length() << " : " << r1name.back() << std::endl; // :1074-1075
    dispatch(smp, r1name, false);
    std::cout << g_tct << " reads loaded" << std::endl;
    if (r2name.length() > 1 && r2name != "none") { // :1099-1121
        dispatch(smp, r2name, true);
        std::cout << g_tct << " reads loaded" << std::endl;
    }

    std::vector<int32_t> gcount((size_t)num_targ), ucount((size_t)num_targ);
    if (kid_sample_counts(smp, gcount.data(), ucount.data(), nullptr) != 0) die(1, kid_last_error());
    std::ofstream out2(wdir + "result.txt");
    for (int i = 0; i < num_targ; i++) out2 << i << "," << gcount[(size_t)i] << "," << ucount[(size_t)i] << "\n";
    out2.close();
    kid_sample_free(smp);
    kid_db_free(db);
    return 0;
}
