python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for a in "--layout M" "--layout M --log2-sectors 28"; do python bench.py --pairs 2000000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $a 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['value'], d['roofline']['lookups_per_s'], d['roofline']['frac'], d['table']['bytes'])"; done
