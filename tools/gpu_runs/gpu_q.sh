#!/bin/bash
# quick check of a device-reader change: per-phase times (R1 then R2, nothing concurrent), outputs against the host reader, the reader's tests
mkdir -p gpurun_out
R=$GRAFT_REPO_ROOT
W=/tmp/kid_v; mkdir -p $W; cd $W
$R/tools/kid_synth db --golden $R/tests/golden/b10 --out $W --den 100 > /dev/null
for i in 0 1 2 3; do $R/tools/kid_synth reads --golden $R/tests/golden/b10 --out $W/fq --sample s$i --pairs 2000000 --first-pair $((i*2000000)) --den 100 > /dev/null; done
KID_GPUS=1 KID_GPU_INGEST=0 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ > $W/host.out 2> $W/host.err; echo "host reader rc=$?"
mkdir -p $W/keep; mv $W/fq/*_result.txt $W/fq/*_reads.txt $W/keep/
KID_SERIAL=1 KID_STATS=1 KID_GPUS=1 KID_GZ_GPU_TIMING=1 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ 2>&1 > /dev/null | grep "on the device" | tail -4 | sed -e 's/.*members; //'
KID_STATS=1 KID_GPUS=1 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ 2>&1 > $W/gpu.out | grep "hits in"
cmp $W/host.out $W/gpu.out && echo "stdout identical"
for f in $W/keep/*; do cmp $f $W/fq/$(basename $f) || echo "$(basename $f) DIFFERS"; done
cd $R
timeout 600 python -m pytest tests/test_gpu_ingest.py -m gpu -q --tb=short -x 2>&1 | tail -3
