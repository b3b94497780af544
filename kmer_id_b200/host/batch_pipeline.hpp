// batch_pipeline.hpp - one host thread keeps dense batches in flight on one or several kid_sample
// objects (kid_classify_dense_async / kid_wait) and hands every finished batch to `done` in stream
// order.  While the GPUs work the reader's threads parse and pack the next batches and this thread
// post-processes the previous one.  Replaces the synchronous per-record call of the reference
// (process_fqgz -> process_qual, newkmer_10nx.cpp:798-801).
//
// Several samples = the shards of ONE logical sample on different GPUs (reads are independent: batches
// are simply dealt round-robin; multi_gpu.hpp merges the shards at sample end).  Two threads may feed
// the same kid_sample at once (R1 and R2) as long as they use different slots: slot_base.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <deque>

#include "../../include/kmer_id.h"
#include "read_reader.hpp"

namespace kidhost {

constexpr int kPipelineSlots = 2; // submissions in flight per (thread, kid_sample)
// reader batches so that every slot can be busy while the parse workers stay busy
inline int pipeline_batches(int n_samples = 1) { return pipeline_depth(kPipelineSlots * n_samples); }

template <class Done>
void classify_stream(kid_sample *const *smps, int n_smps, int slot_base, ReadBatchReader &reader, Done &&done)
{
    struct Flight { ReadBatch *b; kid_sample *smp; };
    std::deque<Flight> inflight;
    size_t submitted = 0;
    auto retire = [&] {
        const Flight f = inflight.front();
        inflight.pop_front();
        if (kid_wait(f.smp, f.b->slot) != 0) {
            fprintf(stderr, "kmer_id_b200: %s\n", kid_last_error());
            exit(1);
        }
        done(*f.b);
        reader.recycle(f.b);
    };
    for (;;) {
        ReadBatch *b = reader.next();
        const bool last = b->last;
        if (b->n) {
            // the slot this batch is about to use was last used n_smps * kPipelineSlots submissions ago
            if ((int)inflight.size() >= n_smps * kPipelineSlots) retire();
            kid_sample *smp = smps[submitted % (size_t)n_smps];
            b->slot = slot_base + (int)((submitted / (size_t)n_smps) % kPipelineSlots);
            submitted++;
            if (kid_classify_dense_async(smp, b->slot, b->codes, b->boff, b->flagbits, b->inv, b->n_inv, b->n, b->taxon) != 0) {
                fprintf(stderr, "kmer_id_b200: %s\n", kid_last_error());
                exit(1);
            }
            inflight.push_back(Flight{ b, smp });
        } else {
            reader.recycle(b);
        }
        if (last) break;
    }
    while (!inflight.empty()) retire();
}

template <class Done>
void classify_stream(kid_sample *smp, ReadBatchReader &reader, Done &&done)
{
    classify_stream(&smp, 1, 0, reader, done);
}

} // namespace kidhost
