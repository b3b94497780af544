#!/bin/bash
# round-2 GPU check Z: warp-parallel inflate + read-ahead of the next sample + cheaper block finder; 5 samples of 2 M pairs
mkdir -p gpurun_out
R=$GRAFT_REPO_ROOT
W=/tmp/kid_v; mkdir -p $W; cd $W
$R/tools/kid_synth db --golden $R/tests/golden/b10 --out $W --den 100 > /dev/null
for i in 0 1 2 3 4; do $R/tools/kid_synth reads --golden $R/tests/golden/b10 --out $W/fq --sample s$i --pairs ${PAIRS:-2000000} --first-pair $((i*2000000)) --den 100 > /dev/null; done
KID_STATS=1 KID_GPUS=1 KID_GPU_INGEST=0 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ > $W/host.out 2> $W/host.err; echo "host reader rc=$?"
grep "reads," $W/host.err
mkdir -p $W/keep; mv $W/fq/*_result.txt $W/fq/*_reads.txt $W/keep/
for P in 32768; do
  echo "== piece $P"
  KID_GZ_GPU_PIECE=$P KID_STATS=1 KID_GPUS=1 KID_GZ_GPU_TIMING=1 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ > $W/gpu.out 2> $W/gpu.err; echo "device reader rc=$?"
  grep -v "^\[nk10\] parse\|cached" $W/gpu.err | tail -n +10 | sed -e 's/.*pieces (/(/' -e 's/.*fastq.gz: /: /' | cut -c1-300
  cmp $W/host.out $W/gpu.out && echo "stdout identical"
  for f in $W/keep/*; do cmp $f $W/fq/$(basename $f) || echo "$(basename $f) DIFFERS"; done
done
echo "== untimed"
KID_STATS=1 KID_GPUS=1 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ 2>&1 >/dev/null | grep "reads,\|total"
cp $W/gpu.err $R/gpurun_out/z_gpu.err
cd $W; rm -f $W/fq/s1_* $W/fq/s2_* $W/fq/s3_* $W/fq/s4_*
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kidz_\(inflate_kernel\|find\|copy\) -c 3 -f -o $R/gpurun_out/prof_r2_ingest env KID_GPUS=1 KID_SERIAL=1 $R/kmer_id_b200/bin/nk10 $W/fq/ > $R/gpurun_out/ncu_z.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $R/gpurun_out/z_launches.csv env KID_GPUS=1 KID_SERIAL=1 $R/kmer_id_b200/bin/nk10 $W/fq/ > /dev/null 2>&1; echo "ncu rc=$?"
