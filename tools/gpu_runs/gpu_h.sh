#!/bin/bash
# round-2 GPU check H: kernel trims, and where the file-level path (nk10 on gz FASTQ) spends its time
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_packed.py tests/test_gpu_parity.py -m gpu -q --tb=short -x > gpurun_out/gputests_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_h.log
tail -n 3 gpurun_out/gputests_h.log
show() { python - "$1" "$2" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]; e=d.get("e2e") or {}
    print("%s value %.1fM e2e %.1fM kernel_ms %.3f pack_ms %.3f G lookups/s %.1f frac %.3f" % (sys.argv[2], d["value"]/1e6, e.get("value",0)/1e6, r["kernel_ms"], r["pack_kernel_ms"], r["lookups_per_s"]/1e9, r["frac"]))
except Exception as ex: print(sys.argv[2], "failed", ex)
P
}
for t in 0 1; do
  KID_TUNE=$t timeout 300 python bench.py --no-cpu-baseline --no-files-e2e > gpurun_out/bench_h_t$t.json 2> gpurun_out/bench_h_t$t.err; show gpurun_out/bench_h_t$t.json "tune $t"
done
R=$GRAFT_REPO_ROOT; W=/tmp/kid_h; mkdir -p $W; cd $W
$R/tools/kid_synth db --golden $R/tests/golden/b10 --out $W --den 1 > /dev/null
for i in 0 1 2 3; do $R/tools/kid_synth reads --golden $R/tests/golden/b10 --out $W/fq --sample s$i --pairs 2000000 --first-pair $((i*2000000)) --den 1 > /dev/null; done
KID_STATS=1 $R/kmer_id_b200/bin/nk10 $W/fq/ > /dev/null 2> $W/warm.err   # writes the probe cache
for cfg in "8 3" "6 3" "6 2" "5 3" "7 1" "8 0" "4 4" "10 3"; do
  set -- $cfg
  KID_STATS=1 KID_GZ_THREADS=$1 KID_PARSE_THREADS=$2 $R/kmer_id_b200/bin/nk10 $W/fq/ > /dev/null 2> $W/run.err
  echo "gz $1 parse $2: $(grep -o 'in [0-9.]* s' $W/run.err | tr '\n' ' ')"
done | tee $R/gpurun_out/h_files_sweep.txt
