#!/bin/bash
# round-2 GPU check W: device-side inflate split into a decode pass and a warp-per-piece copy pass; lane trim in the pack kernel
mkdir -p gpurun_out
R=$GRAFT_REPO_ROOT
W=/tmp/kid_v; mkdir -p $W; cd $W
$R/tools/kid_synth db --golden $R/tests/golden/b10 --out $W --den 100 > /dev/null
for i in 0 1 2; do $R/tools/kid_synth reads --golden $R/tests/golden/b10 --out $W/fq --sample s$i --pairs ${PAIRS:-2000000} --first-pair $((i*2000000)) --den 100 > /dev/null; done
KID_STATS=1 KID_GPUS=1 KID_GPU_INGEST=0 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ > $W/host.out 2> $W/host.err; echo "host reader rc=$?"
grep "reads," $W/host.err
mkdir -p $W/keep; mv $W/fq/*_result.txt $W/fq/*_reads.txt $W/keep/
KID_STATS=1 KID_GPUS=1 KID_GZ_GPU_TIMING=1 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ > $W/gpu.out 2> $W/gpu.err; echo "device reader rc=$?"
grep -v "^\[nk10\] parse\|cached" $W/gpu.err | sed -e 's/.*pieces (/(/' | cut -c1-300
cmp $W/host.out $W/gpu.out && echo "stdout identical"
for f in $W/keep/*; do cmp $f $W/fq/$(basename $f) && echo "$(basename $f) identical"; done
KID_STATS=1 KID_GPUS=1 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ 2>&1 > /dev/null | grep "reads,"
cp $W/gpu.err $R/gpurun_out/w_gpu.err
cd $R
timeout 600 python -m pytest tests/test_nk10_dropin.py tests/test_gpu_packed.py tests/test_gpu_parity.py -m gpu -q --tb=short -x > gpurun_out/gputests_w.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_w.log
tail -n 4 gpurun_out/gputests_w.log
timeout 300 python bench.py --no-cpu-baseline --no-files-e2e --no-e2e > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err
python - <<P
import json
d=json.loads(open("gpurun_out/bench_w.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("value %.1fM ms_per_step %.3f kernel_ms %.3f pack_ms %.3f frac %.3f" % (d["value"]/1e6, d["ms_per_step"], r["kernel_ms"], r["pack_kernel_ms"], r["frac"]))
P
