// batch_pipeline.hpp - one host thread keeps kSlots packed batches in flight on a kid_sample
// (kid_classify_packed_async / kid_wait) and hands every finished batch to `done` in stream order.
// While the GPU works on a batch the reader thread parses and packs the next ones and this thread
// post-processes the previous one.  Replaces the synchronous per-record call of the reference
// (process_fqgz -> process_qual, newkmer_10nx.cpp:798-801).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <deque>

#include "../../include/kmer_id.h"
#include "read_reader.hpp"

namespace kidhost {

constexpr int kPipelineSlots = 2;                       // submissions in flight per kid_sample
inline int pipeline_batches() { return pipeline_depth(kPipelineSlots); } // reader batches: in flight + ready + being filled

template <class Done>
void classify_stream(kid_sample *smp, ReadBatchReader &reader, Done &&done)
{
    std::deque<ReadBatch *> inflight;
    size_t submitted = 0;
    auto retire = [&] {
        ReadBatch *b = inflight.front();
        inflight.pop_front();
        if (b->slot >= 0 && kid_wait(smp, b->slot) != 0) {
            fprintf(stderr, "kmer_id_b200: %s\n", kid_last_error());
            exit(1);
        }
        done(*b);
        reader.recycle(b);
    };
    for (;;) {
        ReadBatch *b = reader.next();
        const bool last = b->last;
        if (b->n) {
            if ((int)inflight.size() >= kPipelineSlots) retire(); // frees the slot this batch is about to use
            b->slot = (int)(submitted++ % kPipelineSlots);
            if (kid_classify_packed_async(smp, b->slot, b->words, 0, b->meta, b->n, b->taxon) != 0) {
                fprintf(stderr, "kmer_id_b200: %s\n", kid_last_error());
                exit(1);
            }
            inflight.push_back(b);
        } else {
            reader.recycle(b);
        }
        if (last) break;
    }
    while (!inflight.empty()) retire();
}

} // namespace kidhost
