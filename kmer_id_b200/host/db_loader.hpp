// db_loader.hpp - text database -> (keys, taxa, parent) arrays for kid_db_build.
// Replaces main():949-989, process_kmergz (:663-712) and process_kmer (:619-661).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace kidhost {

struct ProbeSet {
    std::vector<uint64_t> keys; // forward 60-bit encoding of every 30-window, file order
    std::vector<uint32_t> taxa;
    long long lines_parsed = 0; // the "kmers loaded" figure (:701,:989)
};

// btree file: parent[] as Tree1 leaves it (:101-116, :973-983).  A missing file is not an error in
// the reference.  Returns false (with msg) if an edge names a node outside [0, n_taxa) - the
// reference writes out of bounds there.
bool load_tree(const std::string &path, int n_taxa, std::vector<int32_t> &parent, std::string &msg);

// probes gz: multi-threaded parse (one inflate thread, `threads` parser threads), order preserved.
void load_probes_gz(const std::string &path, ProbeSet &out, unsigned threads = 0);

// one line (without its '\n'); appends windows to keys/taxa; returns true if the line parsed
bool parse_probe_line(const char *line, size_t len, std::vector<uint64_t> &keys,
                      std::vector<uint32_t> &taxa);

} // namespace kidhost
