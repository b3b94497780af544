#!/bin/bash
# round-2 GPU check D3 (2 GPUs): the multi-GPU drop-in tests after the change of the by-sample deal (fixed order, read-ahead),
# then nk10 on 6 samples of 2 M pairs: 1 GPU, 2 GPUs in both modes, outputs compared
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_nk10_dropin.py tests/test_multi_gpu.py -m gpu -q --tb=short -x > gpurun_out/gputests_d3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_d3.log
tail -n 5 gpurun_out/gputests_d3.log
R=$GRAFT_REPO_ROOT
W=/tmp/kid_v; mkdir -p $W; cd $W
$R/tools/kid_synth db --golden $R/tests/golden/b10 --out $W --den 100 > /dev/null
for i in 0 1 2 3 4 5; do $R/tools/kid_synth reads --golden $R/tests/golden/b10 --out $W/fq --sample s$i --pairs 2000000 --first-pair $((i*2000000)) --den 100 > /dev/null; done
KID_STATS=1 KID_GPUS=1 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ > $W/one.out 2> $W/one.err; echo "1 GPU rc=$?"; grep "hits in\|total" $W/one.err
mkdir -p $W/keep; mv $W/fq/*_result.txt $W/fq/*_reads.txt $W/keep/
for mode in reads samples; do
  KID_MULTI_MODE=$mode KID_STATS=1 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ > $W/two.out 2> $W/two.err; echo "2 GPUs, $mode: rc=$?"; grep "hits in\|total" $W/two.err
  cmp $W/one.out $W/two.out && echo "stdout identical"
  for f in $W/keep/*; do cmp $f $W/fq/$(basename $f) || echo "$(basename $f) DIFFERS"; done
done
