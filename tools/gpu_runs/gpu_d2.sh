#!/bin/bash
# round-2 GPU check D2 (2 GPUs): nk10 with the device reader on 2 GPUs (samples dealt to GPUs by default; KID_MULTI_MODE=reads
# puts R1 and R2 of every sample on different GPUs) against the reference, then the torchrun bench at N = 2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_nk10_dropin.py tests/test_golden_ref_case.py tests/test_multi_gpu.py -m gpu -q --tb=short -x > gpurun_out/gputests_d2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_d2.log
tail -n 8 gpurun_out/gputests_d2.log
R=$GRAFT_REPO_ROOT
W=/tmp/kid_v; mkdir -p $W; cd $W
$R/tools/kid_synth db --golden $R/tests/golden/b10 --out $W --den 100 > /dev/null
for i in 0 1 2 3 4 5; do $R/tools/kid_synth reads --golden $R/tests/golden/b10 --out $W/fq --sample s$i --pairs 2000000 --first-pair $((i*2000000)) --den 100 > /dev/null; done
KID_STATS=1 KID_GPUS=1 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ > $W/one.out 2> $W/one.err; echo "1 GPU rc=$?"; grep "hits in\|total" $W/one.err
mkdir -p $W/keep; mv $W/fq/*_result.txt $W/fq/*_reads.txt $W/keep/
for mode in samples reads; do
  KID_MULTI_MODE=$mode KID_STATS=1 timeout 300 $R/kmer_id_b200/bin/nk10 $W/fq/ > $W/two.out 2> $W/two.err; echo "2 GPUs, $mode: rc=$?"; grep "hits in\|total" $W/two.err
  cmp $W/one.out $W/two.out && echo "stdout identical"
  for f in $W/keep/*; do cmp $f $W/fq/$(basename $f) || echo "$(basename $f) DIFFERS"; done
done
cd $R
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_d2_n2.json 2> gpurun_out/bench_d2_n2.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_d2_n2.json").read().strip().splitlines()[-1])
print("N=2 value %.1fM e2e %.1fM" % (d["value"]/1e6, d["e2e"]["value"]/1e6), json.dumps(d.get("files_e2e")))
P
tail -n 3 gpurun_out/bench_d2_n2.err
