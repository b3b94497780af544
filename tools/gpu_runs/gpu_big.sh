#!/bin/bash
# device reader at scale: samples of 12 M and 14 M pairs (3.8 / 4.45 GB of text per file: first run with 32-bit text positions,
# the 14 M sample went to the host reader; second run with 64-bit positions), against the host reader's outputs
R=$GRAFT_REPO_ROOT
W=/tmp/kid_big; mkdir -p $W; cd $W
$R/tools/kid_synth db --golden $R/tests/golden/b10 --out $W --den 100 > /dev/null
$R/tools/kid_synth reads --golden $R/tests/golden/b10 --out $W/fq --sample a12 --pairs 12000000 --den 100 > /dev/null
$R/tools/kid_synth reads --golden $R/tests/golden/b10 --out $W/fq --sample b14 --pairs 14000000 --first-pair 12000000 --den 100 > /dev/null
ls -la $W/fq
KID_STATS=1 KID_GPUS=1 KID_GPU_INGEST=0 timeout 600 $R/kmer_id_b200/bin/nk10 $W/fq/ > $W/host.out 2> $W/host.err; echo "host reader rc=$?"; grep "hits in" $W/host.err
mkdir -p $W/keep; mv $W/fq/*_result.txt $W/fq/*_reads.txt $W/keep/
KID_STATS=1 KID_GPUS=1 timeout 600 $R/kmer_id_b200/bin/nk10 $W/fq/ > $W/gpu.out 2> $W/gpu.err; echo "device reader rc=$?"
grep -v "^\[nk10\] parse\|cached" $W/gpu.err | sed -e 's#/tmp/kid_big/fq/##g' | cut -c1-330
cmp $W/host.out $W/gpu.out && echo "stdout identical"
for f in $W/keep/*; do cmp $f $W/fq/$(basename $f) && echo "$(basename $f) identical"; done
nvidia-smi --query-gpu=memory.used --format=csv
cd $R; timeout 900 python -m pytest tests/test_gpu_ingest.py tests/test_nk10_dropin.py -m gpu -q --tb=short -x 2>&1 | tail -3
