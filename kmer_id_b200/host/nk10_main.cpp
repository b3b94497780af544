// nk10 - GPU drop-in for the reference's `./nk10 <dir/>` (newkmer_10nx.cpp main():915-1054).
//
// Same contract: run from a directory that holds ./bact10/{bData10.txt,btree_10.txt,
// probes10.txt.gz}; argv[1] is the FASTQ directory WITH its trailing slash; for every
// <s>_R1_tr.fastq.gz in it (readdir order) classify R1 then R2 and write <dir><s>_result.txt and
// <dir><s>_reads.txt; same stdout progress lines.  The per-read work happens in
// libkmerid_b200.so (include/kmer_id.h); there is no CPU classification path in this program.
//
// Extras that do not change the contract: all visible GPUs are used (KID_GPUS=<k> limits them,
// KID_DEVICE=<n> pins one; KID_MULTI_MODE=samples deals whole samples to the GPUs instead of splitting
// every sample over them), gz FASTQ files are inflated and
// framed on the GPU (KID_GPU_INGEST=0: on the host), KID_STATS=1 prints phase
// timings to stderr, and the parsed probe list is cached next to probes10.txt.gz as
// probes10.txt.gz.kidcache (stamped with the text file's size and mtime; KID_NO_CACHE=1 disables it).
#include "../../include/kmer_id.h"
#include "batch_pipeline.hpp"
#include "db_loader.hpp"
#include "device_warmup.hpp"
#include "multi_gpu.hpp"
#include "pgz.hpp"
#include "read_reader.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <future>
#include <condition_variable>
#include <iostream>
#include <map>
#include <mutex>
#include <sstream>
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

// ---- gz FASTQ read on the device (kid_fastq_*, include/kmer_id.h): the file's compressed bytes go to one
// GPU, which inflates, frames, trims and classifies them; the host gets the per-read taxa and asks for the
// text of the few reads _reads.txt wants.  KID_GPU_INGEST=0 keeps every file on the host reader.
struct GpuReader {
    kid_fastq *fq = nullptr;
    int32_t *taxon = nullptr; // page-locked
    size_t cap = 0;
    // the same file of the NEXT sample, being read, inflated and framed by a helper thread while this
    // thread still assembles and writes what the current one produced
    std::future<int> ahead;
    std::string ahead_path, ahead_error;
    size_t ahead_reads = 0;
};

GpuReader &gpu_reader(kid_sample *smp, int which)
{
    static std::mutex mu;
    static std::map<std::pair<kid_sample *, int>, GpuReader> readers; // live until exit, like the samples
    std::lock_guard<std::mutex> lk(mu);
    GpuReader &r = readers[{ smp, which }];
    if (!r.fq && kid_fastq_create(kid_sample_db(smp), &r.fq) != 0) die(1, kid_last_error());
    return r;
}

// process_read's bookkeeping (:608-614) replayed over a whole file's taxa: which reads go to _reads.txt (fewer than
// SAVENUM reads of their taxon came before them, in file order), gcount and tct.  Order dependent, so on several
// threads it takes two passes: per-chunk taxon counts, then every chunk starts from the counts before it.
void pick_reads(const int32_t *taxon, size_t n, SampleState &st, std::vector<uint32_t> &pick)
{
    const size_t n_taxa = st.gcount_host.size();
    const unsigned T = n >= ((size_t)1 << 18) ? 4u : 1u;
    std::vector<std::vector<int>> count(T, std::vector<int>(n_taxa, 0));
    std::vector<std::vector<uint32_t>> picks(T);
    std::vector<long long> kept(T, 0);
    auto chunk = [&](unsigned c) { return n * c / T; };
    auto on_chunks = [&](auto &&fn) {
        std::vector<std::thread> th;
        for (unsigned c = 1; c < T; c++) th.emplace_back(fn, c);
        fn(0u);
        for (auto &t : th) t.join();
    };
    on_chunks([&](unsigned c) { // reads per taxon in this chunk
        std::vector<int> &h = count[c];
        long long k = 0;
        for (size_t r = chunk(c); r < chunk(c + 1); r++) {
            const int fin = taxon[r];
            if (fin < 0) continue; // trimmed below 31 bases: the read vanishes (:755)
            h[(size_t)fin]++;
            k++;
        }
        kept[c] = k;
    });
    for (size_t t = 0; t < n_taxa; t++) { // count[c][t] := reads of taxon t before chunk c
        int before = st.gcount_host[t];
        for (unsigned c = 0; c < T; c++) {
            const int here = count[c][t];
            count[c][t] = before;
            before += here;
        }
        st.gcount_host[t] = before;
    }
    on_chunks([&](unsigned c) {
        std::vector<int> &h = count[c];
        std::vector<uint32_t> &out = picks[c];
        for (size_t r = chunk(c); r < chunk(c + 1); r++) {
            const int fin = taxon[r];
            if (fin < 0) continue;
            if (fin > 1 && h[(size_t)fin] < SAVENUM) out.push_back((uint32_t)r);
            h[(size_t)fin]++;
        }
    });
    for (unsigned c = 0; c < T; c++) {
        pick.insert(pick.end(), picks[c].begin(), picks[c].end());
        st.tct += kept[c];
    }
}

// false: the file is one for the host reader (nothing has been counted)
bool run_file_gpu(kid_sample *smp, int which, const std::string &path, const std::string &next_path, SampleState &st,
                  std::ostream *outread, std::vector<SavedRead> *saved, bool stats)
{
    static const bool enabled = !(getenv("KID_GPU_INGEST") && atoi(getenv("KID_GPU_INGEST")) == 0);
    if (!enabled) return false;
    GpuReader &gr = gpu_reader(smp, which);
    const double t_begin = now();
    size_t n = 0;
    int rc;
    std::string error;
    if (gr.ahead.valid() && gr.ahead_path == path) {
        rc = gr.ahead.get();
        n = gr.ahead_reads;
        error = gr.ahead_error;
    } else {
        if (gr.ahead.valid()) gr.ahead.get(); // some other file: that work is lost
        rc = kid_fastq_load_gz_file(gr.fq, path.c_str(), &n);
        if (rc != 0) error = kid_last_error();
    }
    if (rc == KID_EUNSUPPORTED) {
        if (stats) fprintf(stderr, "[nk10] %s\n", error.c_str());
        return false;
    }
    if (rc != 0) die(1, error);
    // the next sample's file: its bytes are read and copied to the device from now on (a second buffer), its
    // kernels start once this file's reads have been fetched (below)
    if (!next_path.empty() && kid_fastq_prefetch_gz_file(gr.fq, next_path.c_str()) != 0) die(1, kid_last_error());
    if (n > gr.cap) {
        kid_host_free(gr.taxon);
        gr.taxon = nullptr;
        gr.cap = n + n / 8 + 1024;
        void *p = nullptr;
        if (kid_host_alloc(&p, sizeof(int32_t) * gr.cap) != 0) die(1, kid_last_error());
        gr.taxon = (int32_t *)p;
    }
    const double t_loaded = now();
    if (kid_fastq_classify(gr.fq, smp, gr.taxon) != 0) die(1, kid_last_error());
    const double t_classified = now();
    // the reads process_read would write (:608-611): fewer than SAVENUM of their taxon came before them
    std::vector<uint32_t> pick;
    pick_reads(gr.taxon, n, st, pick);
    const char *data = nullptr;
    const uint32_t *lens = nullptr;
    if (kid_fastq_fetch(gr.fq, pick.data(), pick.size(), &data, &lens) != 0) die(1, kid_last_error());
    uint64_t st_text = 0, st_pieces = 0, st_again = 0, st_members = 0;
    double ph[8] = { 0 };
    if (stats) kid_fastq_stats(gr.fq, &st_text, &st_pieces, &st_again, &st_members, ph, 8);
    const double t_fetched = now();
    // Everything this file still needs is on the host now (data / lens stay valid until the next fetch, the
    // taxa are in gr.taxon): the next sample's file may take the device buffers.  Loading touches no
    // kid_sample, so it can run under the rest of this sample.
    if (!next_path.empty()) {
        gr.ahead_path = next_path;
        GpuReader *g = &gr;
        gr.ahead = std::async(std::launch::async, [g] {
            const int r = kid_fastq_load_gz_file(g->fq, g->ahead_path.c_str(), &g->ahead_reads);
            g->ahead_error = r != 0 ? kid_last_error() : "";
            return r;
        });
    }
    // ">taxon:header\nbases\n" per read (the bytes of :610), assembled in one buffer
    size_t at = 0, total = 0;
    for (size_t i = 0; i < pick.size(); i++) total += lens[2 * i] + lens[2 * i + 1] + 16;
    std::string text;
    text.reserve(total);
    for (size_t i = 0; i < pick.size(); i++) {
        const int fin = gr.taxon[pick[i]];
        const char *name = data + at, *bases = name + lens[2 * i];
        const size_t nlen = lens[2 * i], blen = lens[2 * i + 1];
        at += nlen + blen;
        char num[16];
        int nd = 0;
        for (int v = fin; v > 0; v /= 10) num[nd++] = (char)('0' + v % 10); // fin > 1
        const size_t rec0 = text.size();
        text.push_back('>');
        while (nd) text.push_back(num[--nd]);
        text.push_back(':');
        text.append(name, nlen);
        text.push_back('\n');
        text.append(bases, blen);
        text.push_back('\n');
        if (saved) {
            SavedRead sr;
            sr.taxon = fin;
            sr.text = text.substr(rec0);
            saved->push_back(std::move(sr));
        }
    }
    if (outread) outread->write(text.data(), (std::streamsize)text.size());
    if (outread) outread->flush();
    if (stats) {
        fprintf(stderr, "[nk10] %s: load (or the wait for it) %.3f classify %.3f pick+fetch %.3f write %.3f s (%zu reads written)\n",
                path.c_str(), t_loaded - t_begin, t_classified - t_loaded, t_fetched - t_classified, now() - t_fetched, pick.size());
        fprintf(stderr, "[nk10] %s on the device: %zu reads, %llu bytes of text in %llu pieces (%llu inflated again), %llu members; "
                        "read %.3f find %.3f inflate %.3f chain %.3f resolve %.3f frame %.3f classify %.3f fetch %.3f s\n",
                path.c_str(), n, (unsigned long long)st_text, (unsigned long long)st_pieces, (unsigned long long)st_again,
                (unsigned long long)st_members, ph[0], ph[1], ph[2], ph[3], ph[4], ph[5], ph[6], ph[7]);
    }
    return true;
}

// Classify one FASTQ file batch by batch on the given shards (one kid_sample per GPU), using their
// slots slot_base, slot_base+1.  direct = R1: append to _reads.txt exactly as process_read :608-614
// does.  Otherwise (R2, running concurrently with R1) keep the first SAVENUM records per taxon in
// stream order; the caller writes those that the reference would have written once R1's per-taxon
// counts are known.
void run_file(kid_sample *const *smps, int n_smps, int slot_base, const std::string &path, const std::string &next_path,
              SampleState &st, std::ostream *outread, std::vector<SavedRead> *saved, unsigned concurrent_files, bool stats)
{
    // the whole file on one GPU: R1 on the first shard, R2 (running next to it) on the second if there is one
    const int which = slot_base ? 1 : 0;
    if (run_file_gpu(smps[which % n_smps], which, path, next_path, st, outread, saved, stats)) return;
    // R1 and R2 (and, with KID_MULTI_MODE=samples, several samples) are inflated at the same time: share the cores
    ReadBatchReader reader(ReadFormat::GzFastq, path, (size_t)1 << 18, kBatchBytes, pipeline_batches(n_smps),
                           default_gz_threads(concurrent_files));
    classify_stream(smps, n_smps, slot_base, reader, [&](const ReadBatch &b) {
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

// One sample (main():1015-1045) on the shards smps[0..n): R1 then R2 (KID_SERIAL) or both at once.
// `log` receives what the reference prints for this sample.  counts() merges the shards.
template <class Counts>
void process_sample(kid_sample *const *smps, int n_smps, const std::string &dname, const std::string &s, const std::string &next,
                    bool serial, unsigned concurrent_samples, std::ostream &log, Counts &&counts, bool stats, GpuSet *stats_set)
{
    // `next`: the sample after this one if it is known already (its files are read ahead), else empty
    const std::string next1 = next.empty() ? std::string() : dname + next + e1, next2 = next.empty() ? std::string() : dname + next + e2;
    const double ts = now();
    SampleState st, st2;
    st.gcount_host.assign((size_t)MAXTAR, 0);
    st2.gcount_host.assign((size_t)MAXTAR, 0);
    const std::string oname2 = dname + s + "_result.txt", trname = dname + s + "_reads.txt";
    log << s << std::endl; // :1022
    std::ofstream outread(trname.c_str(), std::ofstream::out | std::ofstream::trunc);
    long long tct_total = 0;
    if (serial) {
        run_file(smps, n_smps, 0, dname + s + e1, std::string(), st, &outread, nullptr, concurrent_samples, stats);
        log << st.tct << " reads loaded" << std::endl; // :1030
        run_file(smps, n_smps, 0, dname + s + e2, next1, st, &outread, nullptr, concurrent_samples, stats);
        tct_total = st.tct;
    } else {
        // R1 and R2 are inflated, parsed and classified concurrently into the SAME accumulators (gcount and
        // the seen flags are order independent), each thread on its own pair of slots
        std::vector<SavedRead> saved;
        std::thread t2([&] { run_file(smps, n_smps, kPipelineSlots, dname + s + e2, next2, st2, nullptr, &saved, 2 * concurrent_samples, stats); });
        run_file(smps, n_smps, 0, dname + s + e1, next1, st, &outread, nullptr, 2 * concurrent_samples, stats);
        log << st.tct << " reads loaded" << std::endl; // :1030
        t2.join();
        // an R2 read is written iff fewer than SAVENUM reads of its taxon came before it, R1 first
        std::vector<int> seen_r2((size_t)MAXTAR, 0);
        for (const SavedRead &sr : saved) {
            if (st.gcount_host[(size_t)sr.taxon] + seen_r2[(size_t)sr.taxon] < SAVENUM) outread << sr.text << std::flush;
            seen_r2[(size_t)sr.taxon]++;
        }
        tct_total = st.tct + st2.tct;
    }
    std::vector<int32_t> gcount((size_t)MAXTAR), ucount((size_t)MAXTAR);
    counts(gcount.data(), ucount.data());
    log << tct_total << " reads loaded" << std::endl; // :1036
    outread.close();
    std::ofstream out2(oname2);
    for (int i = 0; i < MAXTAR; i++) out2 << i << "," << gcount[(size_t)i] << "," << ucount[(size_t)i] << "\n";
    out2.close();
    if (stats) {
        uint64_t lk = 0, hits = 0;
        if (stats_set) stats_set->counters(lk, hits);
        else kid_sample_counters(smps[0], &lk, &hits, nullptr, nullptr);
        fprintf(stderr, "[nk10] %s: %lld reads, %llu lookups, %llu hits in %.3f s\n", s.c_str(), tct_total,
                (unsigned long long)lk, (unsigned long long)hits, now() - ts);
    }
}
} // namespace

int main(int argc, char *argv[])
{
    const bool stats = getenv("KID_STATS") != nullptr;
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

    GpuSet gpus;
    gpus.devices = GpuSet::pick_devices();
    ProbeSet probes;
    // CUDA context creation overlaps the parse.  (Page-locking the batch buffers here as well was
    // measured and dropped: cudaHostAlloc and the parser's page faults fight over the address-space
    // lock, and the load got 1-3 s slower to save 0.3 s on the first sample.)
    start_device_warmup(gpus.devices[0]);
    const bool cached = load_probes_cached(pname, probes);
    finish_device_warmup();
    const double t1 = now();
    if (!gpus.build(probes.keys.data(), probes.taxa.data(), probes.keys.size(), parent.data(), MAXTAR, 0, msg)) {
        if (msg.find("cannot place") != std::string::npos) {
            std::cout << "out of memory in table " << std::endl; // :258-259
            exit(1);
        }
        die(1, msg);
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

    // KID_SERIAL=1 processes R1 then R2 on one thread, like the reference.
    const bool serial = getenv("KID_SERIAL") != nullptr;
    const int n_gpus = (int)gpus.devices.size();
    // several GPUs: every sample is split over the GPUs (with the device-side reader: R1 on the first, R2 on the
    // second, one exchange at sample end); KID_MULTI_MODE=samples deals whole samples to the GPUs instead (no
    // exchange; GPU g takes samples g, g + n, ...), which is what pays with many samples and more than 2 GPUs
    const char *mm = getenv("KID_MULTI_MODE");
    const bool by_sample = n_gpus > 1 && mm && std::string(mm) == "samples";
    if (!by_sample) {
        // every sample on all GPUs: its batches are dealt to the GPUs, one exchange at sample end
        for (size_t i = 0; i < fnames.size(); i++) {
            const std::string &s = fnames[i];
            if (!gpus.begin(msg)) die(1, msg); // :1017-1019
            process_sample(gpus.samples.data(), n_gpus, dname, s, i + 1 < fnames.size() ? fnames[i + 1] : std::string(), serial, 1, std::cout,
                           [&](int32_t *g, int32_t *u) { if (!gpus.counts(g, u, MAXTAR, msg)) die(1, msg); }, stats, &gpus);
        }
    } else {
        // whole samples dealt to the GPUs: no exchange at all; what each sample prints is held back
        // until every earlier sample has printed, so stdout reads as if they had run one after another
        std::vector<std::string> logs(fnames.size());
        std::vector<char> finished(fnames.size(), 0);
        std::mutex mu;
        std::condition_variable cv;
        size_t next_print = 0;
        std::vector<std::thread> th;
        for (int g = 0; g < n_gpus; g++)
            th.emplace_back([&, g] {
                kid_sample *smp = gpus.samples[(size_t)g];
                // a fixed deal, so that every GPU knows its next sample and reads it ahead
                for (size_t i = (size_t)g; i < fnames.size(); i += (size_t)n_gpus) {
                    if (kid_sample_begin(smp, nullptr) != 0) die(1, kid_last_error());
                    std::ostringstream log;
                    const std::string next = i + (size_t)n_gpus < fnames.size() ? fnames[i + (size_t)n_gpus] : std::string();
                    process_sample(&smp, 1, dname, fnames[i], next, serial, (unsigned)n_gpus, log,
                                   [&](int32_t *gc, int32_t *uc) { if (kid_sample_counts(smp, gc, uc, nullptr) != 0) die(1, kid_last_error()); },
                                   stats, nullptr);
                    {
                        std::lock_guard<std::mutex> lk(mu);
                        logs[i] = log.str();
                        finished[i] = 1;
                        while (next_print < fnames.size() && finished[next_print]) std::cout << logs[next_print++] << std::flush;
                    }
                }
            });
        for (auto &t : th) t.join();
    }
    if (stats)
        fprintf(stderr, "[nk10] %s db %.3f s, build table %.3f s, total %.3f s, %d GPU(s)%s\n", cached ? "cached" : "parse",
                t1 - t0, t2 - t1, now() - t0, n_gpus, n_gpus > 1 ? (by_sample ? ", samples dealt to GPUs" : (gpus.uses_nccl() ? ", reads dealt to GPUs, NCCL all-reduce at sample end" : ", reads dealt to GPUs, host sums at sample end")) : "");
    (void)cached;
    return 0;
}
