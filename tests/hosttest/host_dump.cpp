// host_dump - TEST INFRASTRUCTURE: drives the C++ host parsers of kmer_id_b200/host without a GPU so
// that `pytest -m "not gpu"` can check them (the product hosts link the CUDA library instead of the
// three malloc-backed stubs below).
//   host_dump probes <probes.gz> <signed 0|1> <cap_log2_cells or 0>   -> "lines N\n" then "key taxon" per entry
//   host_dump reads  <gzfastq|fastq|gzfasta|fasta> <file>             -> one record per line: acc \t seq \t qual
//   host_dump packed <gzfastq|fastq|gzfasta|fasta> <file> <flags>     -> per record: acc \t start \t stop \t tlen \t flagged \t decoded
//                                                                       bases (non-ACGT as N), from the reader's packed batches
//   host_dump tree   <tree file> <n_taxa>                             -> parent[] one per line, or "ERR msg"
#include "../../include/kmer_id.h"
#include "../../kmer_id_b200/host/db_loader.hpp"
#include "../../kmer_id_b200/host/read_reader.hpp"

#include <algorithm>
#include <chrono>
#include <ctime>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

extern "C" {
int kid_host_alloc(void **p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? 0 : -3; }
void kid_host_free(void *p) { free(p); }
const char *kid_last_error(void) { return "host_dump stub"; }
}

#include "../../kmer_id_b200/host/gz_lines.hpp"
using namespace kidhost;

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    const std::string cmd = argv[1];
    if (cmd == "probes") {
        ProbeSet ps;
        load_probes_gz(argv[2], ps, argc > 3 && atoi(argv[3]) != 0, 3);
        size_t hidden = 0;
        if (argc > 4 && atoi(argv[4]) > 0) hidden = apply_reference_probe_cap(ps, 16, atoi(argv[4]));
        printf("lines %lld hidden %zu\n", ps.lines_parsed, hidden);
        for (size_t i = 0; i < ps.keys.size(); i++) printf("%llu %u\n", (unsigned long long)ps.keys[i], ps.taxa[i]);
        return 0;
    }
    if (cmd == "reads") {
        const std::string k = argv[2];
        const ReadFormat fmt = k == "gzfastq" ? ReadFormat::GzFastq : k == "fastq" ? ReadFormat::PlainFastq
                             : k == "gzfasta" ? ReadFormat::GzFasta : ReadFormat::PlainFasta;
        ReadBatchReader reader(fmt, argv[3], 7, 4096, 2, 0, BatchMode::Text); // tiny batches: exercise the batch boundaries
        for (;;) {
            ReadBatch *b = reader.next();
            for (size_t r = 0; r < b->n; r++) {
                fwrite(b->names.data() + b->name_off[r], 1, b->name_off[r + 1] - b->name_off[r], stdout);
                fputc('\t', stdout);
                fwrite(b->seq + b->off[r], 1, b->off[r + 1] - b->off[r], stdout);
                fputc('\t', stdout);
                if (b->has_qual) fwrite(b->qual + b->off[r], 1, b->off[r + 1] - b->off[r], stdout);
                fputc('\n', stdout);
            }
            const bool last = b->last;
            reader.recycle(b);
            if (last) break;
        }
        if (reader.open_failed()) printf("OPEN_FAILED\n");
        return 0;
    }
    if (cmd == "packed") {
        const std::string k = argv[2];
        const ReadFormat fmt = k == "gzfastq" ? ReadFormat::GzFastq : k == "fastq" ? ReadFormat::PlainFastq
                             : k == "gzfasta" ? ReadFormat::GzFasta : ReadFormat::PlainFasta;
        ReadBatchReader reader(fmt, argv[3], 7, 4096, 2, 0, BatchMode::Packed, argc > 4 ? (unsigned)atoi(argv[4]) : 0u);
        for (;;) {
            ReadBatch *b = reader.next();
            if (b->n && b->boff[b->n] != b->n_bases) { printf("BAD_END\n"); return 1; }
            for (size_t r = 0; r < b->n; r++) {
                fwrite(b->names.data() + b->name_off[r], 1, b->name_off[r + 1] - b->name_off[r], stdout);
                const uint32_t a = b->boff[r], tlen = b->boff[r + 1] - a, flagged = (b->flagbits[r >> 5] >> (r & 31)) & 1u;
                printf("\t%d\t%d\t%u\t%u\t", (int)b->span[2 * r], (int)b->span[2 * r + 1], tlen, flagged);
                const uint32_t *ia = std::lower_bound(b->inv, b->inv + b->n_inv, a);
                for (uint32_t i = 0; i < tlen; i++) {
                    const uint32_t pos = a + i;
                    const uint32_t c = (b->codes[pos >> 4] >> (30 - 2 * (pos & 15))) & 3u;
                    bool bad = false;
                    if (ia < b->inv + b->n_inv && *ia == pos) { bad = true; ia++; }
                    fputc(bad ? 'N' : "ACGT"[c], stdout);
                }
                fputc('\n', stdout);
            }
            const bool last = b->last;
            reader.recycle(b);
            if (last) break;
        }
        return 0;
    }
    if (cmd == "loadtime") { // <probes.gz> <threads>: timing only
        ProbeSet ps;
        const auto t0 = std::chrono::steady_clock::now();
        timespec c0, c1;
        clock_gettime(CLOCK_THREAD_CPUTIME_ID, &c0);
        load_probes_gz(argv[2], ps, false, (unsigned)atoi(argv[3]));
        clock_gettime(CLOCK_THREAD_CPUTIME_ID, &c1);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("lines %lld keys %zu in %.3f s (calling thread busy %.3f s)\n", ps.lines_parsed, ps.keys.size(), dt,
               (double)(c1.tv_sec - c0.tv_sec) + 1e-9 * (double)(c1.tv_nsec - c0.tv_nsec));
        return 0;
    }
    if (cmd == "readtime") { // <gz fastq>: timing only; the second pass runs on recycled (already touched) buffers
        for (int pass = 0; pass < 2; pass++) {
            const auto t0 = std::chrono::steady_clock::now();
            const bool packed = argc > 3 && atoi(argv[3]) != 0; // the hosts' path: trim + pack on the parse workers
            ReadBatchReader reader(ReadFormat::GzFastq, argv[2], 1u << 18, 48u << 20, packed ? pipeline_depth(2) : 3, 0,
                                   packed ? BatchMode::Packed : BatchMode::Text);
            size_t n = 0, bytes = 0;
            for (;;) {
                ReadBatch *b = reader.next();
                n += b->n;
                bytes += b->off[b->n];
                const bool last = b->last;
                reader.recycle(b);
                if (last) break;
            }
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("pass %d: %zu reads %zu bases in %.3f s = %.2f M reads/s\n", pass, n, bytes, dt, n / dt / 1e6);
        }
        return 0;
    }
    if (cmd == "gunzip") { // <file> <threads> <piece_bytes> [out]: ParallelGunzip alone, min size 0
        auto pg = ParallelGunzip::open(argv[2], (unsigned)atoi(argv[3]), (size_t)atoll(argv[4]), 0);
        if (!pg) { fprintf(stderr, "NOT_APPLICABLE\n"); return 3; }
        FILE *out = argc > 5 ? fopen(argv[5], "wb") : stdout;
        const uint8_t *d;
        size_t n;
        int rc;
        while ((rc = pg->next(d, n)) > 0) fwrite(d, 1, n, out);
        if (out != stdout) fclose(out);
        size_t a, b, c;
        pg->piece_counts(a, b, c);
        fprintf(stderr, "%s delivered %llu pieces %zu again %zu covered %zu\n", rc == 0 ? "END" : "GIVEUP",
                (unsigned long long)pg->delivered(), a, b, c);
        return rc == 0 ? 0 : 4;
    }
    if (cmd == "lines") { // <file> <threads> [out]: GzLineBlocks (parallel with zlib fallback)
        GzLineBlocks src(argv[2], 4u << 20, (unsigned)atoi(argv[3]));
        FILE *out = argc > 4 ? fopen(argv[4], "wb") : stdout;
        std::vector<char> blk;
        while (src.next(blk)) fwrite(blk.data(), 1, blk.size(), out);
        if (out != stdout) fclose(out);
        return 0;
    }
    if (cmd == "tree") {
        std::vector<int32_t> parent;
        std::string msg;
        if (!load_tree(argv[2], atoi(argv[3]), parent, msg)) { printf("ERR %s\n", msg.c_str()); return 0; }
        for (int32_t p : parent) printf("%d\n", p);
        return 0;
    }
    return 2;
}
