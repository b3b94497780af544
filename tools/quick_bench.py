"""Throw-away first-light benchmark: random DB, random (all-miss) reads, device-resident."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import kmer_id_b200 as kid

n_keys = int(float(sys.argv[1])) if len(sys.argv) > 1 else 108_585_519
n_reads = int(float(sys.argv[2])) if len(sys.argv) > 2 else 4_000_000
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
log2s = int(sys.argv[4]) if len(sys.argv) > 4 else 0
L = 150
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
keys = torch.randint(0, 1 << 60, (n_keys,), dtype=torch.int64, device=dev, generator=g)
taxa = torch.randint(2, 5982, (n_keys,), dtype=torch.int32, device=dev, generator=g)
parent = np.ones(5982, np.int32)
t0 = time.time()
db = kid.Database(keys, taxa, parent, flags=flags, log2_sectors=log2s)
torch.cuda.synchronize()
print("build s", time.time() - t0, db.stats())
del keys, taxa
stride = 160
codes = torch.randint(0, 4, (n_reads, stride), dtype=torch.uint8, device=dev, generator=g)
lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
seq = lut[codes.long()].contiguous().view(-1)
seq = torch.cat([seq, torch.zeros(64, dtype=torch.uint8, device=dev)])
qual = torch.full_like(seq, ord("I"))
off = (torch.arange(n_reads + 1, dtype=torch.int64, device=dev) * stride)
# reads are 150 long inside a 160 stride: use explicit ends by making off ragged is not possible,
# so classify 160-base reads (131 k-mers each)
s = kid.Sample(db)
out = torch.empty(n_reads, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
for it in range(3):
    s.begin(st)
    s.classify_device(seq, qual, off, n_reads, out, None, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.begin(st)
e0.record()
reps = 5
for it in range(reps):
    s.classify_device(seq, qual, off, n_reads, out, None, st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
c = s.counters()
lk = c["lookups"] / reps
print(f"classify {ms:.3f} ms  reads/s {n_reads/ms*1e3:.3e}  lookups/s {lk/ms*1e3:.3e}  "
      f"frac_of_204G {lk/ms*1e3*32/6537.6e9:.3f} hits {c['hits']}")
