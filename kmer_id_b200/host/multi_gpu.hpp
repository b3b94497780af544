// multi_gpu.hpp - the C++ hosts on several GPUs of one box: one process, one replica of the probe
// table per GPU (it fits 180 GB many times over), reads dealt to the GPUs batch by batch (they are
// independent, newkmer_10nx.cpp:1029-1031 classifies them one at a time), and ONE exchange per sample:
//   gcount  is additive                      -> ncclAllReduce(sum) over NVLink
//   ucount  is NOT: it is the per-taxon histogram of the OR of all shards' seen
//           bitmaps.  GPU r histograms word range r of that OR with one kernel that reads the peers'
//           bitmaps in place over NVLink (kid_sample_ucount_partial); the ranges are disjoint, so the
//           partial histograms are additive -> a second ncclAllReduce(sum).
// Replaces the per-sample reset / read-out of main() (:1017-1023, :1040-1043) for N GPUs.  Without
// NCCL at build time (or if it fails to initialise) the two small sums are done on the host.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

#include "../../include/kmer_id.h"
#ifdef KID_HAVE_NCCL
#include <nccl.h>
#endif

namespace kidhost {

class GpuSet {
public:
    std::vector<int> devices;
    std::vector<kid_db *> dbs;
    std::vector<kid_sample *> samples; // one per GPU: the shards of the current sample

    // KID_DEVICE=<n> pins one GPU; else KID_GPUS=<k> (default: all visible) uses devices 0..k-1
    static std::vector<int> pick_devices()
    {
        if (const char *e = getenv("KID_DEVICE")) return { atoi(e) };
        int n = 0;
        if (kid_device_count(&n) != 0 || n < 1) return { 0 }; // kid_db_build reports the missing device
        if (const char *e = getenv("KID_GPUS")) n = std::max(1, std::min(n, atoi(e)));
        if (n > 8) n = 8; // the seen-bitmap OR kernel takes up to 16 sources
        std::vector<int> d;
        for (int i = 0; i < n; i++) d.push_back(i);
        return d;
    }

    // one replica per GPU, built concurrently; false + msg on failure
    bool build(const uint64_t *keys, const uint32_t *taxa, size_t n_keys, const int32_t *parent, int n_taxa,
               unsigned flags, std::string &msg)
    {
        const size_t n = devices.size();
        dbs.assign(n, nullptr);
        std::vector<std::string> errs(n);
        std::vector<std::thread> th;
        for (size_t i = 0; i < n; i++)
            th.emplace_back([&, i] {
                if (kid_db_build(keys, taxa, n_keys, 0, parent, n_taxa, devices[i], flags, 0, nullptr, &dbs[i]) != 0)
                    errs[i] = kid_last_error();
            });
        for (auto &t : th) t.join();
        for (size_t i = 0; i < n; i++)
            if (!dbs[i]) { msg = errs[i]; return false; }
        samples.assign(n, nullptr);
        for (size_t i = 0; i < n; i++)
            if (kid_sample_create(dbs[i], &samples[i]) != 0) { msg = kid_last_error(); return false; }
        warm();
        if (n > 1) {
            if (kid_peer_enable(devices.data(), (int)n) != 0) { msg = kid_last_error(); return false; }
#ifdef KID_HAVE_NCCL
            comms_.resize(n);
            // whatever NCCL has to say while it starts ("NCCL version ..." with NCCL_DEBUG=VERSION goes to
            // stdout) must not end up among the reference's stdout lines: stdout is stderr for that long
            fflush(stdout);
            const int saved = dup(1);
            if (saved >= 0) dup2(2, 1);
            if (getenv("KID_NO_NCCL") == nullptr && ncclCommInitAll(comms_.data(), (int)n, devices.data()) == ncclSuccess)
                nccl_ = true;
            else
                comms_.clear();
            fflush(stdout);
            if (saved >= 0) { dup2(saved, 1); close(saved); }
#endif
        }
        return true;
    }

    bool uses_nccl() const { return nccl_; }

    // One tiny dense batch through every slot of every shard: loads the kernels (CUDA loads a module on its
    // first launch), creates the slot streams and their first buffers, so that the first sample does not pay
    // for it.  The accumulators it touches are cleared by the kid_sample_begin that starts every sample.
    void warm()
    {
        static const char bases[] = "ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT"; // 64
        void *pin = nullptr;
        if (kid_host_alloc(&pin, 4096) != 0) return;
        uint32_t *codes = (uint32_t *)pin, *boff = codes + 64, *flagbits = codes + 80, *inv = codes + 96;
        int32_t *taxon = (int32_t *)(codes + 128);
        const uint64_t off[2] = { 0, 64 };
        size_t ni = 0;
        uint32_t nb = 0;
        if (kid_pack_reads_dense((const uint8_t *)bases, nullptr, off, 1, 0, 0, codes, 64, boff, flagbits, 0, inv, 16, &ni, nullptr, &nb) == 0)
            for (kid_sample *s : samples) {
                for (int slot = 0; slot < KID_MAX_SLOTS; slot++) kid_classify_dense_async(s, slot, codes, boff, flagbits, inv, ni, 1, taxon);
                for (int slot = 0; slot < KID_MAX_SLOTS; slot++) kid_wait(s, slot);
            }
        kid_host_free(pin);
    }

    bool begin(std::string &msg)
    {
        for (kid_sample *s : samples)
            if (kid_sample_begin(s, nullptr) != 0) { msg = kid_last_error(); return false; }
        return true;
    }

    // sample end: gcount / ucount of the whole sample (all shards) on the host
    bool counts(int32_t *gcount, int32_t *ucount, int n_taxa, std::string &msg)
    {
        const int n = (int)samples.size();
        if (n == 1) {
            if (kid_sample_counts(samples[0], gcount, ucount, nullptr) != 0) { msg = kid_last_error(); return false; }
            return true;
        }
        auto sync_all = [&] {
            for (int d : devices)
                if (kid_device_sync(d) != 0) { msg = kid_last_error(); return false; }
            return true;
        };
        if (!sync_all()) return false; // every shard's seen bits are written
        for (int r = 0; r < n; r++)
            if (kid_sample_ucount_partial(samples[(size_t)r], samples.data(), n, r, n, nullptr) != 0) { msg = kid_last_error(); return false; }
        bool reduced = false;
#ifdef KID_HAVE_NCCL
        if (nccl_) {
            bool ok = ncclGroupStart() == ncclSuccess;
            for (int r = 0; r < n && ok; r++) {
                int32_t *g = nullptr, *u = nullptr;
                kid_sample_gcount_device(samples[(size_t)r], &g);
                kid_sample_ucount_device(samples[(size_t)r], &u);
                // stream 0 of each GPU: ordered after that GPU's partial-histogram kernel
                ok = ncclAllReduce(g, g, (size_t)n_taxa, ncclInt32, ncclSum, comms_[(size_t)r], nullptr) == ncclSuccess &&
                     ncclAllReduce(u, u, (size_t)n_taxa, ncclInt32, ncclSum, comms_[(size_t)r], nullptr) == ncclSuccess;
            }
            ok = ncclGroupEnd() == ncclSuccess && ok;
            if (!ok) { msg = "NCCL all-reduce failed"; return false; }
            if (kid_sample_read_counts(samples[0], gcount, ucount, nullptr) != 0) { msg = kid_last_error(); return false; }
            reduced = true;
        }
#endif
        if (!reduced) { // two sums of n_taxa integers on the host
            std::vector<int32_t> g((size_t)n_taxa), u((size_t)n_taxa);
            for (int t = 0; t < n_taxa; t++) gcount[t] = ucount[t] = 0;
            for (int r = 0; r < n; r++) {
                if (kid_sample_read_counts(samples[(size_t)r], g.data(), u.data(), nullptr) != 0) { msg = kid_last_error(); return false; }
                for (int t = 0; t < n_taxa; t++) { gcount[t] += g[(size_t)t]; ucount[t] += u[(size_t)t]; }
            }
        }
        return sync_all(); // peers are done reading every bitmap before the next begin() clears it
    }

    void counters(uint64_t &lookups, uint64_t &hits)
    {
        lookups = hits = 0;
        for (kid_sample *s : samples) {
            uint64_t a = 0, b = 0;
            kid_sample_counters(s, &a, &b, nullptr, nullptr);
            lookups += a;
            hits += b;
        }
    }

    ~GpuSet()
    {
#ifdef KID_HAVE_NCCL
        for (ncclComm_t c : comms_) ncclCommDestroy(c);
#endif
        for (kid_sample *s : samples) kid_sample_free(s);
        for (kid_db *d : dbs) kid_db_free(d);
    }

private:
    bool nccl_ = false;
#ifdef KID_HAVE_NCCL
    std::vector<ncclComm_t> comms_;
#endif
};

} // namespace kidhost
