"""One-off measurements of the other BASELINE.json configs (parity-test cases, not bench lines):
   config 1: mitochondria DB (22 008 398 probes, 17 227 taxa), 1 M 150-bp pairs
   config 4: 10x bact10 synthetic DB (1.086 G probes), 250-bp pairs
   python tools/bench_configs.py mito | x10"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import kmer_id_b200 as kid
from tools import synthlib

which = sys.argv[1]
dev = torch.device("cuda:0")
if which == "mito":
    parent, prefix = synthlib.load_taxonomy(os.path.join(ROOT, "tests/golden/mito"), 1, 1,
                                            tree="mitochondria_tree.txt", refkey="mitochondria_refkey.txt")
    L, pairs, seeds = 150, 1_000_000, (11, 22)
else:
    parent, prefix = synthlib.load_taxonomy(os.path.join(ROOT, "tests/golden/b10"), 10, 1)
    L, pairs, seeds = 250, 4_000_000, (10, 25)
wl = synthlib.Workload(parent, prefix, read_len=L, seed_db=seeds[0], seed_reads=seeds[1])
print("probes", wl.n_probes, "taxa", wl.n_taxa)
dk = torch.empty(wl.n_probes, dtype=torch.int64, device=dev)
dt = torch.empty(wl.n_probes, dtype=torch.int32, device=dev)
wl.db_device(0, dk, dt)
torch.cuda.synchronize()
t0 = time.time()
db = kid.Database(dk, dt, parent)
torch.cuda.synchronize()
print("build s", round(time.time() - t0, 3), db.stats())
del dk, dt
n = 2 * pairs
dseq = torch.empty(n * L + 64, dtype=torch.uint8, device=dev)
dqual = torch.empty(n * L + 64, dtype=torch.uint8, device=dev)
wl.reads_device(0, 0, n, dseq, dqual)
doff = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
out = torch.empty(n, dtype=torch.int32, device=dev)
s = kid.Sample(db)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    s.begin(st); s.classify_device(dseq, dqual, doff, n, out, None, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
s.begin(st)
e0.record()
for _ in range(reps):
    s.classify_device(dseq, dqual, doff, n, out, None, st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
c = s.counters()
lk = c["lookups"] / reps
g, u = s.counts()
print(json.dumps({"config": which, "probes": wl.n_probes, "read_len": L, "pairs": pairs, "kernel_ms": ms,
                  "pairs_per_s": pairs / ms * 1e3, "lookups_per_s": lk / ms * 1e3,
                  "frac_of_32B_hbm_roofline": lk / ms * 1e3 * 32 / 6537.6e9, "table": db.stats(),
                  "classified_fraction": float(g[2:].sum() / max(1, g.sum())), "hit_fraction": c["hits"] / c["lookups"]}))
