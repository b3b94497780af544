#!/bin/bash
# end of round 2, what the driver runs: the GPU suite, smoke(), the default bench line; then nk10's own phase timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -m gpu -q --tb=short -x > gpurun_out/gputests_p3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_p3.log
tail -n 4 gpurun_out/gputests_p3.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_p3.json 2> gpurun_out/bench_p3.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open("gpurun_out/bench_p3.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("value %.1fM e2e %.1fM ms_per_step %.3f kernel_ms %.3f pack_ms %.3f frac %.3f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], r["kernel_ms"], r["pack_kernel_ms"], r["frac"]))
f=d["files_e2e"]; print("files_e2e %.1fM" % (f["value"]/1e6), f["sample_s"], f["reader"], "whole run %.2fM" % (f["whole_run_pairs_per_s"]/1e6))
P
R=$GRAFT_REPO_ROOT
W=/tmp/kid_v; mkdir -p $W; cd $W
$R/tools/kid_synth db --golden $R/tests/golden/b10 --out $W --den 100 > /dev/null
for i in 0 1 2 3 4; do $R/tools/kid_synth reads --golden $R/tests/golden/b10 --out $W/fq --sample s$i --pairs 2000000 --first-pair $((i*2000000)) --den 100 > /dev/null; done
KID_STATS=1 KID_GPUS=1 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ 2>&1 > /dev/null | grep "hits in\|load (or" | sed -e 's#.*fastq.gz: #  #' | tee $R/gpurun_out/p3_nk10_phases.txt
