// nk10 - GPU drop-in for the reference's `./nk10 <dir/>` (newkmer_10nx.cpp main():915-1054).
//
// Same contract: run from a directory that holds ./bact10/{bData10.txt,btree_10.txt,
// probes10.txt.gz}; argv[1] is the FASTQ directory WITH its trailing slash; for every
// <s>_R1_tr.fastq.gz in it (readdir order) classify R1 then R2 and write <dir><s>_result.txt and
// <dir><s>_reads.txt; same stdout progress lines.  The per-read work happens in
// libkmerid_b200.so (include/kmer_id.h); there is no CPU classification path in this program.
//
// Extras that do not change the contract: KID_DEVICE=<n> picks the GPU, KID_STATS=1 prints phase
// timings to stderr, and the parsed probe list is cached next to probes10.txt.gz as
// probes10.txt.gz.kidcache (stamped with the text file's size and mtime; KID_NO_CACHE=1 disables it).
#include "../../include/kmer_id.h"
#include "batch_pipeline.hpp"
#include "db_loader.hpp"
#include "device_warmup.hpp"
#include "pgz.hpp"
#include "read_reader.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

using namespace kidhost;

namespace {
const int MAXTAR = 5982;                    // newkmer_10nx.cpp:45
const int SAVENUM = 12;                     // :48
const std::string e1 = "_R1_tr.fastq.gz";   // :29
const std::string e2 = "_R2_tr.fastq.gz";   // :30
const std::string iname = "./bact10/bData10.txt";      // :67
const std::string pname = "./bact10/probes10.txt.gz";  // :68
const std::string tname = "./bact10/btree_10.txt";     // :69

double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

[[noreturn]] void die(int code, const std::string &msg)
{
    std::cerr << "nk10: " << msg << std::endl;
    exit(code);
}

constexpr size_t kBatchBytes = (size_t)48 << 20; // bases per batch (one asynchronous submission)

struct SampleState {
    std::vector<int> gcount_host; // running copy, only to decide the first SAVENUM reads per taxon
    long long tct = 0;
};

struct SavedRead { // a _reads.txt record of the R2 file, held back until R1 is complete
    int taxon;
    std::string text; // ">taxon:acc\nbases\n"
};

// Classify one FASTQ file batch by batch.  direct = R1: append to _reads.txt exactly as process_read
// :608-614 does.  Otherwise (R2, running concurrently with R1) keep the first SAVENUM records per
// taxon in stream order; main() writes those that the reference would have written once R1's
// per-taxon counts are known.
void run_file(kid_sample *smp, const std::string &path, SampleState &st, std::ofstream *outread,
              std::vector<SavedRead> *saved)
{
    // R1 and R2 are inflated at the same time unless KID_SERIAL is set: share the cores
    ReadBatchReader reader(ReadFormat::GzFastq, path, (size_t)1 << 18, kBatchBytes, pipeline_batches(),
                           default_gz_threads(getenv("KID_SERIAL") ? 1 : 2));
    classify_stream(smp, reader, [&](const ReadBatch &b) {
        for (size_t r = 0; r < b.n; r++) {
            const int fin = b.taxon[r];
            if (fin < 0) continue; // trimmed below 31 bases: the read vanishes (:755)
            if (fin > 1 && st.gcount_host[(size_t)fin] < SAVENUM) {
                const char *name = b.names.data() + b.name_off[r];
                const size_t nlen = b.name_off[r + 1] - b.name_off[r];
                const char *bases = (const char *)b.seq + b.off[r] + b.span[2 * r];
                const size_t blen = b.span[2 * r + 1] - b.span[2 * r] + 1;
                if (outread) {
                    *outread << ">" << fin << ":";
                    outread->write(name, (std::streamsize)nlen);
                    *outread << std::endl;
                    outread->write(bases, (std::streamsize)blen);
                    *outread << std::endl;
                } else if (saved) {
                    SavedRead sr;
                    sr.taxon = fin;
                    sr.text = ">" + std::to_string(fin) + ":" + std::string(name, nlen) + "\n" +
                              std::string(bases, blen) + "\n";
                    saved->push_back(std::move(sr));
                }
            }
            st.gcount_host[(size_t)fin]++;
            st.tct++;
        }
    });
}
} // namespace

int main(int argc, char *argv[])
{
    const bool stats = getenv("KID_STATS") != nullptr;
    const int device = getenv("KID_DEVICE") ? atoi(getenv("KID_DEVICE")) : 0;
    std::string dname = argc > 1 ? argv[1] : "/mnt/dmb/Mark_backup/J/"; // :933-942
    const double t0 = now();

    // strain list: only feeds the (dead) alignment branch, but its absence is reported (:951-971)
    {
        std::ifstream fin(iname);
        if (!fin) std::cout << "narin " << iname << std::endl;
    }
    std::vector<int32_t> parent;
    std::string msg;
    if (!load_tree(tname, MAXTAR, parent, msg)) die(1, msg);
    std::cout << "tree loaded" << std::endl; // :984

    ProbeSet probes;
    // CUDA context creation overlaps the parse.  (Page-locking the batch buffers here as well was
    // measured and dropped: cudaHostAlloc and the parser's page faults fight over the address-space
    // lock, and the load got 1-3 s slower to save 0.3 s on the first sample.)
    start_device_warmup(device);
    const bool cached = load_probes_cached(pname, probes);
    finish_device_warmup();
    const double t1 = now();
    kid_db *db = nullptr;
    if (kid_db_build(probes.keys.data(), probes.taxa.data(), probes.keys.size(), 0, parent.data(), MAXTAR,
                     device, 0, 0, nullptr, &db) != 0) {
        if (std::string(kid_last_error()).find("cannot place") != std::string::npos) {
            std::cout << "out of memory in table " << std::endl; // :258-259
            exit(1);
        }
        die(1, kid_last_error());
    }
    std::cout << probes.lines_parsed << " kmers loaded" << std::endl; // :989
    { ProbeSet().keys.swap(probes.keys); std::vector<uint32_t>().swap(probes.taxa); }
    const double t2 = now();

    // all samples in the directory, readdir order (:992-1014)
    std::vector<std::string> fnames;
    std::cout << dname << std::endl;
    if (DIR *dir = opendir(dname.c_str())) {
        while (struct dirent *ent = readdir(dir)) {
            const std::string name1 = ent->d_name;
            const size_t pos = name1.find(e1);
            if (pos != std::string::npos) fnames.push_back(name1.substr(0, pos));
        }
        closedir(dir);
    } else {
        std::cout << "hosed" << std::endl;
        perror("");
        return EXIT_FAILURE;
    }

    // R1 and R2 are inflated, parsed and classified concurrently, each into its own kid_sample
    // (gcount and the seen flags are order independent); the library merges the two at sample end.
    // KID_SERIAL=1 processes R1 then R2 on one sample, like the reference.
    const bool serial = getenv("KID_SERIAL") != nullptr;
    kid_sample *smp = nullptr, *smp2 = nullptr;
    if (kid_sample_create(db, &smp) != 0) die(1, kid_last_error());
    if (!serial && kid_sample_create(db, &smp2) != 0) die(1, kid_last_error());
    std::vector<int32_t> gcount((size_t)MAXTAR), ucount((size_t)MAXTAR);
    for (const std::string &s : fnames) {
        const double ts = now();
        if (kid_sample_begin(smp, nullptr) != 0) die(1, kid_last_error()); // :1017-1019
        if (smp2 && kid_sample_begin(smp2, nullptr) != 0) die(1, kid_last_error());
        SampleState st, st2;
        st.gcount_host.assign((size_t)MAXTAR, 0);
        st2.gcount_host.assign((size_t)MAXTAR, 0);
        const std::string oname2 = dname + s + "_result.txt", trname = dname + s + "_reads.txt";
        std::cout << s << std::endl; // :1022
        std::ofstream outread(trname.c_str(), std::ofstream::out | std::ofstream::trunc);
        long long tct_total = 0;
        if (serial) {
            run_file(smp, dname + s + e1, st, &outread, nullptr);
            std::cout << st.tct << " reads loaded" << std::endl; // :1030
            run_file(smp, dname + s + e2, st, &outread, nullptr);
            tct_total = st.tct;
            if (kid_sample_counts(smp, gcount.data(), ucount.data(), nullptr) != 0) die(1, kid_last_error());
        } else {
            std::vector<SavedRead> saved;
            std::thread t2([&] { run_file(smp2, dname + s + e2, st2, nullptr, &saved); });
            run_file(smp, dname + s + e1, st, &outread, nullptr);
            std::cout << st.tct << " reads loaded" << std::endl; // :1030
            t2.join();
            // an R2 read is written iff fewer than SAVENUM reads of its taxon came before it, R1 first
            std::vector<int> seen_r2((size_t)MAXTAR, 0);
            for (const SavedRead &sr : saved) {
                if (st.gcount_host[(size_t)sr.taxon] + seen_r2[(size_t)sr.taxon] < SAVENUM) outread << sr.text << std::flush;
                seen_r2[(size_t)sr.taxon]++;
            }
            tct_total = st.tct + st2.tct;
            kid_sample *both[2] = { smp, smp2 };
            if (kid_samples_counts(both, 2, gcount.data(), ucount.data(), nullptr) != 0) die(1, kid_last_error());
        }
        std::cout << tct_total << " reads loaded" << std::endl; // :1036
        outread.close();
        std::ofstream out2(oname2);
        for (int i = 0; i < MAXTAR; i++) out2 << i << "," << gcount[(size_t)i] << "," << ucount[(size_t)i] << "\n";
        out2.close();
        st.tct = tct_total;
        if (stats) {
            uint64_t lk = 0, hits = 0, rd = 0;
            kid_sample_counters(smp, &lk, &hits, &rd, nullptr);
            if (smp2) {
                uint64_t lk2 = 0, hits2 = 0;
                kid_sample_counters(smp2, &lk2, &hits2, &rd, nullptr);
                lk += lk2;
                hits += hits2;
            }
            fprintf(stderr, "[nk10] %s: %lld reads, %llu lookups, %llu hits in %.3f s\n", s.c_str(), st.tct,
                    (unsigned long long)lk, (unsigned long long)hits, now() - ts);
        }
    }
    if (stats)
        fprintf(stderr, "[nk10] %s db %.3f s, build table %.3f s, total %.3f s\n", cached ? "cached" : "parse", t1 - t0,
                t2 - t1, now() - t0);
    (void)cached;
    kid_sample_free(smp);
    kid_sample_free(smp2);
    kid_db_free(db);
    return 0;
}
