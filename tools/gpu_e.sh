#!/bin/bash
mkdir -p gpurun_out /tmp/w && cd /tmp/w
R=$GRAFT_REPO_ROOT
env | grep -i nccl
$R/tools/kid_synth db --golden $R/tests/golden/b10 --out /tmp/w --den 1000 > /dev/null
$R/tools/kid_synth reads --golden $R/tests/golden/b10 --out /tmp/w/fq --sample a --pairs 20000 --den 1000 > /dev/null
KID_STATS=1 $R/kmer_id_b200/bin/nk10 /tmp/w/fq/ > $R/gpurun_out/e_stdout.txt 2> $R/gpurun_out/e_stderr.txt; echo rc=$?
head -c 600 $R/gpurun_out/e_stdout.txt; echo ----; head -c 1500 $R/gpurun_out/e_stderr.txt
