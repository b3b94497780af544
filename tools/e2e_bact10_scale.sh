#!/bin/bash
# Full-scale drop-in check on the GPU box: bact10-scale synthetic DB in the reference's own text format
# (108 585 519 probe lines), 2 M synthetic pairs, OUR nk10 (parse, then cached) and the UNMODIFIED
# reference nk10 on the same files; outputs must be byte-identical.  Writes gpurun_out/e2e_bact10.log
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=${1:-/tmp/kid_e2e}
PAIRS=${2:-2000000}
LOG=$ROOT/gpurun_out/e2e_bact10.log
mkdir -p "$W" "$ROOT/gpurun_out"; : > "$LOG"
t() { date +%s.%N; }
bc() { python3 -c "import sys; print(round(eval(sys.stdin.read()), 2))"; }
say() { echo "$@" | tee -a "$LOG"; }
say "host: $(nproc) cores, $(free -g | awk '/Mem/{print $2}') GB RAM"
s=$(t); "$ROOT/tools/kid_synth" db --golden "$ROOT/tests/golden/b10" --out "$W" --den 1 >> "$LOG"; say "generate DB text: $(echo "$(t) - $s" | bc) s, $(du -sh "$W/bact10/probes10.txt.gz" | cut -f1)"
s=$(t); "$ROOT/tools/kid_synth" reads --golden "$ROOT/tests/golden/b10" --out "$W/fq" --sample big --pairs "$PAIRS" --den 1 >> "$LOG"; say "generate reads: $(echo "$(t) - $s" | bc) s"
cd "$W"
# kid_synth writes many gzip members in parallel; real files are one member (history runs across the
# whole file), so recompress with gzip itself before timing anything
s=$(t)
for f in bact10/probes10.txt.gz fq/big_R1_tr.fastq.gz fq/big_R2_tr.fastq.gz; do ( zcat "$f" | gzip -1 > "$f.one" && mv "$f.one" "$f" ) & done; wait
ln -s big_R1_tr.fastq.gz fq/big2_R1_tr.fastq.gz; ln -s big_R2_tr.fastq.gz fq/big2_R2_tr.fastq.gz  # a second sample: steady-state time per sample
say "recompress as single-member gzip -1: $(echo "$(t) - $s" | bc) s, $(du -sh bact10/probes10.txt.gz | cut -f1)"
s=$(t); KID_NO_CACHE=1 KID_GZ_THREADS=1 KID_STATS=1 "$ROOT/kmer_id_b200/bin/nk10" "$W/fq/" > ours0.out 2> ours0.err; rc=$?; say "OURS (parse text DB, zlib inflate on one thread = round-1 path): rc=$rc wall $(echo "$(t) - $s" | bc) s"; cat ours0.err >> "$LOG"
cp fq/big_result.txt zlib_result.txt
s=$(t); KID_STATS=1 "$ROOT/kmer_id_b200/bin/nk10" "$W/fq/" > ours1.out 2> ours1.err; rc=$?; say "OURS (parse text DB, multi-threaded inflate): rc=$rc wall $(echo "$(t) - $s" | bc) s"; cat ours1.err >> "$LOG"
cp fq/big_result.txt ours_result.txt; cp fq/big_reads.txt ours_reads.txt
s=$(t); KID_STATS=1 "$ROOT/kmer_id_b200/bin/nk10" "$W/fq/" > ours2.out 2> ours2.err; rc=$?; say "OURS (cached DB):     rc=$rc wall $(echo "$(t) - $s" | bc) s"; cat ours2.err >> "$LOG"
cmp fq/big_result.txt ours_result.txt && cmp zlib_result.txt ours_result.txt && say "cached run == parsed run == zlib-inflate run"
if [ -n "${SKIP_REF:-}" ]; then say "reference run skipped (SKIP_REF)"; rm -rf "$W"; exit 0; fi
s=$(t); "$ROOT/oracle/_ref/nk10" "$W/fq/" 2> ref.err | while IFS= read -r line; do echo "$(t) $line"; done > ref.stamped; say "REFERENCE nk10 (unmodified, 1 thread): wall $(echo "$(t) - $s" | bc) s"
cut -d' ' -f2- ref.stamped > ref.out
awk -v s0="$s" '{printf "  +%.1f s  %s\n", $1 - s0, substr($0, index($0,$2))}' ref.stamped | tee -a "$LOG"
cmp ref.out ours1.out && say "stdout identical"
cmp fq/big_result.txt ours_result.txt && say "_result.txt identical to the reference's ($(wc -l < ours_result.txt) lines)"
cmp fq/big_reads.txt ours_reads.txt && say "_reads.txt identical to the reference's ($(wc -l < ours_reads.txt) lines)"
awk -F, '{g+=$2;u+=$3} END{print "reads counted", g, "distinct hit k-mers", u}' ours_result.txt | tee -a "$LOG"
rm -rf "$W"
