#!/bin/bash
# round-2 GPU check V: gz FASTQ read on the device (kid_ingest.cu) - drop-in tests, then nk10 on 2 M-pair samples
mkdir -p gpurun_out
R=$GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_nk10_dropin.py tests/test_golden_ref_case.py -m gpu -q --tb=short -x > gpurun_out/gputests_v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_v.log
tail -n 15 gpurun_out/gputests_v.log
W=/tmp/kid_v; mkdir -p $W; cd $W
$R/tools/kid_synth db --golden $R/tests/golden/b10 --out $W --den 100 > /dev/null
for i in 0 1 2; do $R/tools/kid_synth reads --golden $R/tests/golden/b10 --out $W/fq --sample s$i --pairs ${PAIRS:-2000000} --first-pair $((i*2000000)) --den 100 > /dev/null; done
ls -la $W/fq | head
KID_STATS=1 KID_GPUS=1 KID_GPU_INGEST=0 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ > $W/host.out 2> $W/host.err; echo "host reader rc=$?"
grep "reads," $W/host.err
mkdir -p $W/keep; mv $W/fq/*_result.txt $W/fq/*_reads.txt $W/keep/
KID_STATS=1 KID_GPUS=1 KID_GZ_GPU_TIMING=1 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ > $W/gpu.out 2> $W/gpu.err; echo "device reader rc=$?"
grep -v "^\[nk10\] parse\|cached" $W/gpu.err | cut -c1-400
cmp $W/host.out $W/gpu.out && echo "stdout identical"
for f in $W/keep/*; do cmp $f $W/fq/$(basename $f) && echo "$(basename $f) identical"; done
cp $W/gpu.err $R/gpurun_out/v_gpu.err; cp $W/host.err $R/gpurun_out/v_host.err
