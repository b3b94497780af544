#!/bin/bash
# ncu --set full of the device reader's three big kernels as committed, on one 2 M-read file
mkdir -p gpurun_out
R=$GRAFT_REPO_ROOT
W=/tmp/kid_v; mkdir -p $W; cd $W
$R/tools/kid_synth db --golden $R/tests/golden/b10 --out $W --den 100 > /dev/null
$R/tools/kid_synth reads --golden $R/tests/golden/b10 --out $W/fq --sample s0 --pairs 2000000 --den 100 > /dev/null
KID_GPUS=1 KID_SERIAL=1 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ > /dev/null 2>&1; echo "plain rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kidz_\(inflate_kernel\|find\|copy\) -c 3 -f -o $R/gpurun_out/prof_r2_ingest2 env KID_GPUS=1 KID_SERIAL=1 $R/kmer_id_b200/bin/nk10 $W/fq/ > $R/gpurun_out/ncu_n2.log 2>&1; echo "ncu rc=$?"
