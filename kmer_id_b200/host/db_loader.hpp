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
// target_signed: the target column is extracted into an `int` (kmer_read_m3.cpp:847,875,
// kmer_read_vf6.cpp) instead of an `unsigned int` (newkmer_10nx.cpp:670,697) - only odd lines differ.
void load_probes_gz(const std::string &path, ProbeSet &out, bool target_signed = false, unsigned threads = 0);

// one line (without its '\n'); appends windows to keys/taxa; returns true if the line parsed
bool parse_probe_line(const char *line, size_t len, std::vector<uint64_t> &keys,
                      std::vector<uint32_t> &taxa, bool target_signed = false);

// Binary cache of a parsed probe file: `<path>.kidcache` holds the (keys, taxa)
// arrays and the parsed-line count, stamped with the size and mtime of the text file it came from.
// load_probes_cached() uses it when the stamp matches and otherwise parses `path` and (unless
// KID_NO_CACHE is set or the directory is read-only) writes it.  Returns true if the cache was used.
bool load_probes_cached(const std::string &path, ProbeSet &out, bool target_signed = false);

// kmer_read_m3.cpp:42,232 - getHash gives up after MAXREPROBE = 16 probes while add_kmer (:235-264)
// does not: a key whose FIRST inserted copy sits deeper than 16 probes in the reference's
// 2^30-cell triangular-probing table is invisible.  Replays the reference's insertion order on an
// occupancy bitmap and zeroes the taxon of every copy of such keys.  Returns how many keys were hidden.
size_t apply_reference_probe_cap(ProbeSet &probes, int cap = 16, int log2_cells = 30);

} // namespace kidhost
