python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
for t in 0 4 5 1; do echo -n "tune $t: "; KID_TUNE=$t python bench.py --pairs 2000000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['value'], d['roofline']['lookups_per_s'], d['roofline']['frac'])"; done
