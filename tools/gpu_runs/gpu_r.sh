#!/bin/bash
# where the read phase of the device reader spends its time when files are loaded ahead
R=$GRAFT_REPO_ROOT
W=/tmp/kid_v; mkdir -p $W; cd $W
nproc; free -g | head -2; cat /sys/fs/cgroup/cpu.max 2>/dev/null
$R/tools/kid_synth db --golden $R/tests/golden/b10 --out $W --den 100 > /dev/null
for i in 0 1 2 3 4; do $R/tools/kid_synth reads --golden $R/tests/golden/b10 --out $W/fq --sample s$i --pairs 2000000 --first-pair $((i*2000000)) --den 100 > /dev/null; done
KID_STATS=1 KID_GPUS=1 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ 2>&1 >/dev/null | grep "reads,"
echo "== timing"
KID_STATS=1 KID_GPUS=1 KID_GZ_GPU_TIMING=1 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ 2>&1 >/dev/null | grep "reads,\|kid_fastq" | sed -e 's#/tmp/kid_v/fq/##'
exit 0
KID_SYNC_BLOCKING=1 KID_STATS=1 KID_GPUS=1 KID_GZ_GPU_TIMING=1 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ 2>&1 >/dev/null | grep "reads,\|kid_fastq" | sed -e 's#/tmp/kid_v/fq/##'
