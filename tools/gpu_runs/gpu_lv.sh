#!/bin/bash
# device reader on gzip level 6 / 9 files (longer matches, fewer tokens than the benchmark's level 1), single member
R=$GRAFT_REPO_ROOT
W=/tmp/kid_lv; mkdir -p $W; cd $W
$R/tools/kid_synth db --golden $R/tests/golden/b10 --out $W --den 100 > /dev/null
for lv in 6 9; do
  rm -rf $W/fq; 
  for i in 0 1 2; do $R/tools/kid_synth reads --golden $R/tests/golden/b10 --out $W/fq --sample s$i --pairs 2000000 --first-pair $((i*2000000)) --den 100 --level $lv --members 100000000 > /dev/null; done
  echo "== level $lv"; ls -la $W/fq | head -4
  KID_STATS=1 KID_GPUS=1 KID_GPU_INGEST=0 timeout 600 $R/kmer_id_b200/bin/nk10 $W/fq/ > $W/host.out 2> $W/host.err; echo "host reader rc=$?"; grep "hits in" $W/host.err
  mkdir -p $W/keep; rm -f $W/keep/*; mv $W/fq/*_result.txt $W/fq/*_reads.txt $W/keep/
  KID_STATS=1 KID_GPUS=1 KID_GZ_GPU_TIMING=1 timeout 600 $R/kmer_id_b200/bin/nk10 $W/fq/ > $W/gpu.out 2> $W/gpu.err; echo "device reader rc=$?"
  grep "hits in\|on the device\|left to" $W/gpu.err | sed -e 's#/tmp/kid_lv/fq/##g' -e 's/.*bytes of text in/  /' | cut -c1-250
  cmp $W/host.out $W/gpu.out && echo "stdout identical"
  for f in $W/keep/*; do cmp $f $W/fq/$(basename $f) || echo "$(basename $f) DIFFERS"; done
done
