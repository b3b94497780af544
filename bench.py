#!/usr/bin/env python
"""bench.py - the nk10 read-classification hot path on B200, one process per GPU.

    python bench.py --gpus N --steps K --warmup W              (ours; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU path)
    python bench.py --config mito | x10                              (BASELINE.json configs[1] / [4])

Default workload (BASELINE.json configs[2], the largest single-GPU configuration): synthetic
bact10-scale probe database (108 585 519 probes = sum of the shipped refkey10.txt counts, b10
taxonomy) and 10 M synthetic 150-bp read pairs per GPU (tools/synth/kid_synth.h).  A *step* is one
whole sample: kid_sample_begin -> classify every read of the batch -> sample end (ucount histogram,
counts to the host; for N > 1 the cross-rank gcount sum / seen-bitmap OR).  A pair is two
independently classified reads.

  value  pairs/s with the TEXT batch (bases + qualities) already resident in HBM: kid_pack_kernel
         (trim + 2-bit pack) + kid_classify3_kernel per step, CUDA events on the launching stream
  e2e    pairs/s through kid_classify_dense_host from pinned HOST buffers holding the dense batch a
         parser produces with kid_pack_reads_dense (H2D + expand + scan + D2H inside the region);
         e2e_packed: the word-aligned packed format (kid_classify_packed_host); e2e_text: the same
         through kid_classify_host from bases + qualities (what round 1 reported as e2e)
  roofline.achieved = lookups/launch x 32 B / mean kid_classify3_kernel time (32 B = one DRAM sector
         per lookup); peak = MEASURED_PEAKS.json hbm_gbs
  cpu_baseline: the CPU oracle (a port, oracle/kid_oracle.c) single-threaded on a bounded sample;
         its per-read taxa are compared with the GPU's for the same reads (parity_checked_reads)
  files_e2e    gz FASTQ files on disk -> _result.txt through kmer_id_b200/bin/nk10 (N = 1 only)

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN_B10 = os.path.join(ROOT, "tests", "golden", "b10")
GOLDEN_MITO = os.path.join(ROOT, "tests", "golden", "mito")
METRIC = "paired_reads_per_sec_classified"
UNIT = "pairs/s"
SECTOR_BYTES = 32

CONFIGS = {
    "bact10": dict(label="BASELINE.json configs[2]: bact10-scale synthetic probe DB + 10M synthetic 150bp pairs per GPU",
                   golden=GOLDEN_B10, tree="btree_10.txt", refkey="refkey10.txt", num=1, read_len=150,
                   pairs=10_000_000, seeds=(10, 21), taxonomy="b10"),
    "mito": dict(label="BASELINE.json configs[1]: mitochondria DB (kmer_read_m3 path) + 1M synthetic 150bp pairs per GPU",
                 golden=GOLDEN_MITO, tree="mitochondria_tree.txt", refkey="mitochondria_refkey.txt", num=1,
                 read_len=150, pairs=1_000_000, seeds=(11, 22), taxonomy="mitochondria"),
    "x10": dict(label="BASELINE.json configs[4]: 10x bact10 synthetic probe DB + 4M synthetic 250bp pairs per GPU",
                golden=GOLDEN_B10, tree="btree_10.txt", refkey="refkey10.txt", num=10, read_len=250,
                pairs=4_000_000, seeds=(10, 25), taxonomy="b10"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="bact10", choices=sorted(CONFIGS))
    ap.add_argument("--pairs", type=int, default=0, help="read pairs per GPU per step (0 = the config's)")
    ap.add_argument("--db-den", type=int, default=1, help="probe DB = refkey counts x num / den")
    ap.add_argument("--cpu-baseline-reads", type=int, default=400_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-files-e2e", action="store_true",
                    help="skip the gz-files-on-disk -> _result.txt run of the nk10 drop-in")
    ap.add_argument("--layout", default="M", choices=["M", "K"], help="table layout (M = minimizer, default)")
    ap.add_argument("--log2-sectors", type=int, default=0, help="table size override (0 = library default)")
    ap.add_argument("--parity-reads", type=int, default=-1,
                    help="N > 1: reads of rank 0 checked against the CPU oracle over the full probe list "
                         "(-1 = 100000, but 0 for --config x10 whose oracle needs ~105 GB of host RAM and minutes)")
    ap.add_argument("--chunk-reads", type=int, default=0, help="reads per H2D chunk of the host entry points (0 = library default)")
    ap.add_argument("--ref-seconds", type=float, default=80.0,
                    help="CPU seconds the reference arm may spend classifying (all steps together)")
    ap.add_argument("--ref-den", type=int, default=1,
                    help="reference arm: probe DB = refkey counts / den (1 = the benchmark's full DB; "
                         "falls back to 100, and says so, when disk or RAM are short)")
    a = ap.parse_args()
    a.cfg = CONFIGS[a.config]
    if a.pairs <= 0:
        a.pairs = a.cfg["pairs"]
    a.read_len = a.cfg["read_len"]
    if a.parity_reads < 0:
        a.parity_reads = 0 if a.config == "x10" else 100_000
    return a


def workload_config(args, n_probes, den=None, pairs=None):
    den = args.db_den if den is None else den
    pairs = args.pairs if pairs is None else pairs
    L = args.read_len
    return {
        "workload": args.cfg["label"] + ("" if den == 1 else " - probe DB sampled 1/%d" % den),
        "db_probes": int(n_probes),
        "db": "%s taxonomy, refkey probe counts x %d / %d, random canonical 30-mers (seed %d)"
              % (args.cfg["taxonomy"], args.cfg["num"], den, args.cfg["seeds"][0]),
        "pairs_per_gpu_per_step": int(pairs),
        "read_len": L,
        "reads": "70%% stitched from lineage probes, 0.5%% subs, 0.1%% N, 20%% low-quality tails (seed %d)" % args.cfg["seeds"][1],
        "sharding": "reads sharded across ranks, table replicated",
        "l2": "inputs_larger_than_l2 (%.1f GB of reads per step, GB-scale probe table: see table.bytes)"
              % (pairs * 2 * L * 2 / 1e9),
    }


def make_workload(args, den=None):
    from tools import synthlib
    c = args.cfg
    parent, prefix = synthlib.load_taxonomy(c["golden"], c["num"], args.db_den if den is None else den,
                                            tree=c["tree"], refkey=c["refkey"])
    return parent, synthlib.Workload(parent, prefix, read_len=c["read_len"], seed_db=c["seeds"][0], seed_reads=c["seeds"][1])


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons while a timed region runs (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1)] or \
               [r for _, r in self.rows]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if len(r) > 3 + i and r[3 + i] == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(rows)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def live_gather_ceiling(table_bytes):
    """Independent random 32-byte reads per second on a buffer of the table's size
    (tools/micro/gather_bench.cu, its own process, after the timed region): the practical ceiling
    of any one-sector-per-lookup table.  (None, None) if the tool is not built."""
    exe = os.path.join(ROOT, "tools", "micro", "gather_bench")
    if not os.path.exists(exe):
        return None, None
    mib = max(1024, int(table_bytes) >> 20)
    try:
        r = subprocess.run([exe, str(mib), "4", "256", "8"], capture_output=True, text=True, timeout=120)
        for line in r.stdout.splitlines():
            if "G accesses/s" in line:
                return float(line.split("ms")[1].split("G accesses/s")[0]), f"live: gather_bench {mib} MiB"
    except Exception:
        pass
    return None, None


def traffic_per_lookup():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def bind_to_gpu_numa(local_rank: int):
    """Pin this rank (and therefore its first-touch pinned host buffers) to the NUMA node its GPU hangs
    off, so that with N ranks the H2D streams do not all cross the inter-socket link.  Plumbing only;
    silently does nothing when sysfs does not say."""
    try:
        import torch
        bdf = torch.cuda.get_device_properties(local_rank).pci_bus_id
    except Exception:
        bdf = None
    try:
        if not bdf:
            out = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                 capture_output=True, text=True, timeout=10).stdout.strip()
            bdf = out
        bdf = bdf.lower()
        if len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = set(os.sched_getaffinity(0)) & set(cpus)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        return None
    return None


# ------------------------------------------------------------------------------------ ours
def pack_dense_on_host(kid, np, hseq, hqual, n_reads, read_len, torch):
    """kid_pack_reads_dense (the parser-side writer of dense batches) over the whole batch on ONE host thread,
    into pinned memory; returns (DenseBatch, seconds).  Untimed setup: a parser does this per record."""
    nbytes = n_reads * read_len
    pin = lambda n: torch.empty(n, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
    dense = kid.DenseBatch(n_reads, nbytes, codes=pin(int(kid.lib.kid_dense_bound(nbytes))), boff=pin(n_reads + 1),
                           flagbits=pin(n_reads // 32 + 2), inv=pin(max(1024, nbytes // 100)))
    seq, qual = hseq.numpy(), hqual.numpy()
    step = 1 << 16
    t0 = time.perf_counter()
    for r0 in range(0, n_reads, step):
        n = min(step, n_reads - r0)
        off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(read_len))
        dense.append(seq[r0 * read_len:(r0 + n) * read_len], qual[r0 * read_len:(r0 + n) * read_len], off)
    return dense, time.perf_counter() - t0


def pack_on_host(kid, np, hseq, hqual, n_reads, read_len, hwords, hmeta):
    """kid_pack_reads (the parser-side writer of packed batches) over the whole batch on ONE host
    thread, slice by slice so that the words of consecutive slices follow each other; returns
    (n_words, seconds).  Untimed setup: a parser does this while it still has each record in cache."""
    step = 1 << 16
    nw_total = 0
    t0 = time.perf_counter()
    seq, qual = hseq.numpy(), hqual.numpy()
    words, meta = hwords.numpy().view(np.uint32), hmeta.numpy().view(np.uint32)
    for r0 in range(0, n_reads, step):
        n = min(step, n_reads - r0)
        off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(read_len))
        w, _ = kid.pack_reads(seq[r0 * read_len:], qual[r0 * read_len:], off, word0=nw_total,
                              words=words[nw_total:], meta=meta[2 * r0:])
        nw_total += int(w.size)
    return nw_total, time.perf_counter() - t0


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import kmer_id_b200 as kid  # raises if the CUDA library is not built: no fallback
    from kmer_id_b200 import multi_gpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    READ_LEN = args.read_len
    parent, wl = make_workload(args)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- database: generated on the device, built by the library's own kernels
    dk = torch.empty(wl.n_probes, dtype=torch.int64, device=dev)
    dt = torch.empty(wl.n_probes, dtype=torch.int32, device=dev)
    wl.db_device(local, dk, dt, stream=stream)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    db = kid.Database(dk, dt, parent, device=local, stream=stream, log2_sectors=args.log2_sectors,
                      flags=kid.KID_DB_LAYOUT_KEYHASH if args.layout == "K" else 0)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    st = db.stats()
    keep_keys = world == 1 and not args.no_cpu_baseline or world > 1 and rank == 0 and args.parity_reads > 0
    hk = dk.cpu().numpy().view(np.uint64) if keep_keys else None  # the oracle checks below want the probe list
    ht = dt.cpu().numpy().view(np.uint32) if keep_keys else None
    del dk, dt
    torch.cuda.empty_cache()

    # ---- this rank's reads, resident in HBM as text (bases + qualities)
    n_reads = 2 * args.pairs
    nbytes = n_reads * READ_LEN
    dseq = torch.empty(nbytes + 64, dtype=torch.uint8, device=dev)
    dqual = torch.empty(nbytes + 64, dtype=torch.uint8, device=dev)
    wl.reads_device(local, rank * n_reads, n_reads, dseq, dqual, stream=stream)
    doff = torch.arange(n_reads + 1, dtype=torch.int64, device=dev) * READ_LEN
    dout = torch.empty(n_reads, dtype=torch.int32, device=dev)
    sample = kid.Sample(db)
    if args.chunk_reads:
        sample.set_chunk_reads(args.chunk_reads)
    engine, transport = multi_gpu.make_engine(sample, stream, prefer_peer=os.environ.get("KID_PEER", "1") != "0")
    torch.cuda.synchronize()

    def step_device():
        sample.begin(stream)
        sample.classify_device(dseq, dqual, doff, n_reads, dout, None, stream)
        return multi_gpu.finish(engine)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        gcount, ucount = step_device()
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.3)
    l_before = kid.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tc0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        gcount, ucount = step_device()
    e1.record()
    barrier()
    gpu_launches = kid.kernel_launches() - l_before
    ms = e0.elapsed_time(e1)
    counters = sample.counters(stream)  # of the last step
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * args.pairs / (ms_per_step / 1e3)
    reads_kept = int((dout >= 0).sum().item())  # this rank's reads that passed the length rule (:755)
    if world > 1:  # every read any rank kept is in the reduced gcount exactly once
        t = torch.tensor([reads_kept], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        reads_kept = int(t.item())
    assert int(gcount.astype(np.int64).sum()) == reads_kept, "gcount does not add up to the reads classified"

    # ---- the dominant kernel alone: packed batch resident in HBM, CUDA events around each launch
    cap = int(kid.lib.kid_pack_bound(n_reads, nbytes))
    k_ms, pack_ms = [], []
    if args.layout == "M":
        dwords = torch.empty(cap, dtype=torch.int32, device=dev)
        dmeta = torch.empty(2 * (n_reads + 1), dtype=torch.int32, device=dev)
        for i in range(args.warmup + args.steps):
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            sample.begin(stream)
            a.record()
            db.pack_device(dseq, dqual, doff, nbytes, n_reads, dwords, cap, dmeta, None, stream)
            b.record()
            sample.classify_packed_device(dwords, dmeta, n_reads, dout, stream)
            c.record()
            torch.cuda.synchronize()
            if i >= args.warmup:
                pack_ms.append(a.elapsed_time(b))
                k_ms.append(b.elapsed_time(c))
        g3, _ = sample.counts(stream)
        if world == 1:
            assert np.array_equal(g3, gcount), "packed-device path and text path disagree"
        del dwords, dmeta
    else:
        for i in range(args.steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sample.begin(stream)
            a.record()
            sample.classify_device(dseq, dqual, doff, n_reads, dout, None, stream)
            b.record()
            torch.cuda.synchronize()
            k_ms.append(a.elapsed_time(b))
    kernel_ms = statistics.mean(k_ms)
    gpu_taxa = dout.cpu().numpy()

    # ---- end to end through the host-buffer entry points
    e2e = e2e_text = e2e_packed = host_pack = None
    if not args.no_e2e:
        hseq = torch.empty(nbytes + 64, dtype=torch.uint8, pin_memory=True)
        hqual = torch.empty(nbytes + 64, dtype=torch.uint8, pin_memory=True)
        hseq.copy_(dseq)
        hqual.copy_(dqual)
        hout = torch.empty(n_reads, dtype=torch.int32, pin_memory=True)
        torch.cuda.synchronize()

        def timed(fn):
            hout.fill_(-7)
            for _ in range(max(1, min(args.warmup, 2))):
                g2, u2 = fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                g2, u2 = fn()
            barrier()
            dt_s = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt_s], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt_s = float(t.item())
            h2d, d2h = sample.transfer_bytes()
            assert np.array_equal(g2, gcount) and np.array_equal(u2, ucount), "host and device paths disagree"
            assert np.array_equal(hout.numpy(), gpu_taxa)
            return {"value": world * args.pairs * args.steps / dt_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h + 8 * db.n_taxa),
                    "ms_per_step": dt_s / args.steps * 1e3}

        if args.layout == "M":
            hwords = torch.empty(cap, dtype=torch.int32, pin_memory=True)
            hmeta = torch.empty(2 * (n_reads + 1), dtype=torch.int32, pin_memory=True)
            n_words, pack_s = pack_on_host(kid, np, hseq, hqual, n_reads, READ_LEN, hwords, hmeta)
            host_pack = {"reads_per_s_per_core": n_reads / pack_s, "text_GB_per_s_per_core": 2 * nbytes / pack_s / 1e9,
                         "bytes_per_read": (4 * n_words + 8 * (n_reads + 1)) / n_reads,
                         "what": "kid_pack_reads (process_qual trim + 2-bit pack) on one host thread, outside the timed region"}

            def step_packed():
                sample.begin(stream)
                sample.classify_packed_host(hwords, hmeta, n_reads, hout)
                return multi_gpu.finish(engine)

            e2e_packed = timed(step_packed)
            e2e_packed["input"] = "packed batch in pinned host memory (kid_pack_reads format), kid_classify_packed_host"
            del hwords, hmeta
            dense, dense_s = pack_dense_on_host(kid, np, hseq, hqual, n_reads, READ_LEN, torch)
            host_pack["dense_reads_per_s_per_core"] = n_reads / dense_s
            host_pack["dense_bytes_per_read"] = dense.wire_bytes / n_reads

            def step_dense():
                sample.begin(stream)
                sample.classify_dense_host(dense, hout)
                return multi_gpu.finish(engine)

            e2e = timed(step_dense)
            e2e["input"] = "dense batch in pinned host memory (kid_pack_reads_dense format), kid_classify_dense_host"

        def step_text():
            sample.begin(stream)
            sample.classify_host(hseq, hqual, hoff, n_reads, hout, None)
            return multi_gpu.finish(engine)

        hoff = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(READ_LEN))
        e2e_text = timed(step_text)
        e2e_text["input"] = "bases + qualities + offsets in pinned host memory, kid_classify_host"
        if e2e is None:
            e2e = e2e_text
    clk = clocks.stop(tc0, None)

    # ---- N > 1: rank 0 checks its first reads against the CPU oracle (N = 1 does it in cpu_baseline)
    parity_n = 0
    if world > 1 and rank == 0 and args.parity_reads > 0:
        parity_n = oracle_check(args, wl, hk, ht, parent, gpu_taxa, args.parity_reads)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    lookups = counters["lookups"]
    peak, peak_src = measured_peak()
    achieved = lookups * SECTOR_BYTES / (kernel_ms / 1e3) / 1e9
    tr = traffic_per_lookup()
    ceiling, ceiling_src = live_gather_ceiling(st["table_bytes"]) if world == 1 else (None, None)
    if ceiling is None:
        ceiling, ceiling_src = (tr or {}).get("gather_ceiling_gsectors_s"), "profiles/traffic.json"
    roofline = {"bound": "hbm", "kernel": "kid_classify3_kernel" if args.layout == "M" else "kid_classify_kernel",
                "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "traffic": (tr["dram_bytes_per_lookup"] * lookups if tr and args.config == "bact10" else None),
                "traffic_source": (("constant: %.2f DRAM bytes per lookup (ncu dram__bytes_read+write of %s) x this run's lookups; "
                                    "not re-measured in this run") % (tr["dram_bytes_per_lookup"], tr.get("source", "profiles/traffic.json"))
                                   if tr and args.config == "bact10" else None),
                "algorithmic_bytes_per_lookup": SECTOR_BYTES, "lookups_per_launch": lookups,
                "kernel_ms": kernel_ms, "lookups_per_s": lookups / (kernel_ms / 1e3),
                "pack_kernel_ms": statistics.mean(pack_ms) if pack_ms else None,
                "pack_kernel_text_GB_per_s": (2 * nbytes / (statistics.mean(pack_ms) / 1e3) / 1e9 if pack_ms else None),
                "random_sector_gather_ceiling_gsectors_s": ceiling, "gather_ceiling_source": ceiling_src}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 keys / int32 counts", "data": "synthetic",
        "config": workload_config(args, wl.n_probes), "clocks": clk, "e2e": e2e, "e2e_packed": e2e_packed, "e2e_text": e2e_text,
        "host_pack": host_pack,
        "gpu_launches": int(gpu_launches), "roofline": roofline,
        "lookups_per_s_whole_step": world * lookups / (ms_per_step / 1e3),
        "table": {"layout": args.layout, "bytes": st["table_bytes"], "distinct_keys": st["n_distinct"], "displaced": st["n_displaced"],
                  "build_s": build_s},
        "sample_end": {"peer": "one fused OR+histogram kernel over NVLink-mapped peer bitmaps + 2 small all-reduces",
                       "nccl": "all-to-all of bitmap slices + OR kernel + histogram kernel + 2 small all-reduces"}[transport]
                      if world > 1 else "single GPU: histogram kernel",
        "numa_node_rank0": numa,
        "hit_fraction": counters["hits"] / max(1, lookups),
        "classified_fraction": float((gcount[2:].sum()) / max(1, gcount.sum())),
        "reads_kept_all_ranks": int(reads_kept),
    }
    if parity_n:
        out["parity_checked_reads"] = parity_n

    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_port(args, wl, hk, ht, parent, gpu_taxa)
    if not args.no_files_e2e and args.config == "bact10":
        out["files_e2e"] = files_e2e(args, world)  # the C++ host on all `world` GPUs, in its own process
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def files_e2e(args, n_gpus=1):
    """SURVEY.md 8(d) "end-to-end (gz on disk -> _result.txt)": the shipped drop-in
    kmer_id_b200/bin/nk10 on files, after the timed region, in its own process.  The probe DB in the
    reference's text format at the benchmark's scale and three samples of gz FASTQ (generator's
    multi-member gzip); times are nk10's own (KID_STATS).  None if the binaries are not built."""
    import re
    import shutil
    import tempfile
    nk10 = os.path.join(ROOT, "kmer_id_b200", "bin", "nk10")
    synth = os.path.join(ROOT, "tools", "kid_synth")
    if not (os.path.exists(nk10) and os.path.exists(synth)):
        return None
    pairs = min(args.pairs, 2_000_000)
    work = tempfile.mkdtemp(prefix="kid_files_")
    try:
        fq = os.path.join(work, "fq")
        subprocess.run([synth, "db", "--golden", GOLDEN_B10, "--out", work, "--den", str(args.db_den)], check=True,
                       stdout=subprocess.DEVNULL)
        n_samples = 5
        for i in range(n_samples):
            subprocess.run([synth, "reads", "--golden", GOLDEN_B10, "--out", fq, "--sample", "s%d" % i, "--pairs",
                            str(pairs), "--first-pair", str(i * pairs), "--den", str(args.db_den)], check=True,
                           stdout=subprocess.DEVNULL)
        t0 = time.perf_counter()
        r = subprocess.run([nk10, fq + "/"], cwd=work, capture_output=True, text=True,
                           env=dict(os.environ, KID_STATS="1", KID_NO_CACHE="1", KID_GPUS=str(n_gpus)))
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {"error": "nk10 exited %d: %s" % (r.returncode, r.stderr[-300:])}
        per_sample = [float(x) for x in re.findall(r"\[nk10\] s\d+: \d+ reads, \d+ lookups, \d+ hits in ([0-9.]+) s", r.stderr)]
        on_device_files = len(re.findall(r"on the device: \d+ reads", r.stderr))
        on_device = on_device_files == 2 * n_samples
        m = re.search(r"\[nk10\] parse db ([0-9.]+) s, build table ([0-9.]+) s, total ([0-9.]+) s", r.stderr)
        med = statistics.median(per_sample)
        phases = None
        try:  # the device-side reader's own phase times per file (seconds, median over the files after the first sample)
            rows = [dict((k, float(v)) for k, v in re.findall(r"(read|find|inflate|chain|resolve|frame|classify|fetch) (-?[0-9.]+)", ln))
                    for ln in r.stderr.splitlines() if "on the device:" in ln][2:]
            if rows:
                phases = {k: statistics.median(row[k] for row in rows if k in row) for k in rows[0]}
        except Exception:
            phases = None
        return {"value": pairs / med, "unit": UNIT, "reader_phase_s": phases,
                "reader_phase_note": "host-clock seconds per file between the reader's synchronisation points, R1 and R2 "
                                     "overlapping; find is launched with inflate and shows up there; most of it runs ahead, "
                                     "under the previous sample",
                "what": "kmer_id_b200/bin/nk10: gz FASTQ on disk -> _result.txt/_reads.txt, MEDIAN of %d samples in one "
                        "process (%s); probe DB parsed from gz text"
                        % (n_samples, "compressed bytes to the GPU: inflate, line framing, trim, pack and classify in kernels, "
                                      "read-ahead of the next sample" if on_device else
                                      "inflate on all host cores, parse + trim + pack on the host, classify on the GPU"),
                "reader": "device" if on_device else "host", "files_on_device": on_device_files,
                "n_gpus": n_gpus, "mode": (re.findall(r"GPU\(s\)(.*)", r.stderr) or [""])[0].strip(", "),
                "pairs_per_sample": pairs, "sample_s": per_sample,
                "first_sample_pairs_per_s": pairs / per_sample[0], "best_sample_pairs_per_s": pairs / min(per_sample),
                "db_parse_s": float(m.group(1)),
                "table_build_s": float(m.group(2)), "process_total_s": float(m.group(3)), "wall_s": wall,
                "whole_run_pairs_per_s": n_samples * pairs / wall, "host_cores": os.cpu_count()}
    finally:
        shutil.rmtree(work, ignore_errors=True)


def oracle_check(args, wl, keys, taxa, parent, gpu_taxa, n, timing=None):
    """first n reads of rank 0's batch through the CPU oracle over the FULL probe table; aborts the
    bench on the first read whose taxon differs from what the timed GPU steps produced"""
    import numpy as np
    from oracle import kor
    n = min(n, gpu_taxa.size)
    odb = kor.OracleDB(wl.n_taxa)
    odb.set_parents(parent)
    odb.add_keys(keys, taxa)
    seq, qual = wl.reads_host(0, n)
    off = wl.offsets(n)
    osamp = kor.OracleSample(odb)
    t0 = time.perf_counter()
    fin, _ = osamp.classify(seq, qual, off)
    dt_s = time.perf_counter() - t0
    if not np.array_equal(fin, gpu_taxa[:n]):
        bad = int(np.flatnonzero(fin != gpu_taxa[:n])[0])
        raise SystemExit("bench.py: GPU and oracle disagree on read %d of the benchmark batch: %d vs %d"
                         % (bad, int(gpu_taxa[bad]), int(fin[bad])))
    if timing is not None:
        timing.update(seconds=dt_s, lookups=osamp.lookups)
    return int(n)


def cpu_baseline_port(args, wl, keys, taxa, parent, gpu_taxa):
    """The CPU oracle (kind = "port"), one thread, on the first cpu_baseline_reads reads of the same
    workload against the FULL probe table (so its cache behaviour is the real one).  Its per-read
    taxa double as a parity check of what the timed GPU steps computed for those reads."""
    tm = {}
    n = oracle_check(args, wl, keys, taxa, parent, gpu_taxa, min(args.cpu_baseline_reads, 2 * args.pairs), tm)
    return {"value": (n / 2) / tm["seconds"], "unit": UNIT, "cores": 1, "kind": "port", "parity_checked_reads": int(n),
            "sample": "first %d reads of rank 0's batch, full %d-probe table, oracle/kid_oracle.c" % (n, keys.size),
            "lookups_per_s": tm["lookups"] / tm["seconds"], "seconds": tm["seconds"]}


# ------------------------------------------------------------------------------------ reference
def run_reference(args):
    """The reference's own CPU implementation: oracle/_ref/nk10 (unmodified newkmer_10nx.cpp,
    single-threaded - it has no threading) on the benchmark's probe DB in its own text format and
    (warmup+steps) samples of P pairs each, one nk10 process; a step is one sample, timed from its
    name line to its second "reads loaded" line on stdout.  P is bounded by --ref-seconds; the
    `config` printed is the one actually run."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import shutil
    import tempfile
    nk10 = os.path.join(ROOT, "oracle", "_ref", "nk10")
    synth = os.path.join(ROOT, "tools", "kid_synth")
    n_samples = args.warmup + args.steps
    if args.config != "bact10" or not (os.path.exists(nk10) and os.path.exists(synth)):
        # mito: the m3 reader's CLI differs; x10: the reference cannot load it (newkmer_10nx.cpp:256-260)
        return run_reference_port(args)
    pairs = max(2000, int(args.ref_seconds * 16000 / max(1, n_samples)))  # ~32 k reads/s single thread
    den = args.ref_den
    note = ""
    if den == 1:
        free_disk = shutil.disk_usage(tempfile.gettempdir()).free
        try:
            avail_ram = int(next(l for l in open("/proc/meminfo") if l.startswith("MemAvailable")).split()[1]) * 1024
        except Exception:
            avail_ram = 0
        if free_disk < 6 << 30 or avail_ram < 40 << 30:
            den, note = 100, " (full DB needs 6 GB of disk and 40 GB of RAM: %.0f / %.0f GB free)" % (free_disk / 2**30, avail_ram / 2**30)
    work = tempfile.mkdtemp(prefix="kid_ref_")
    fq = os.path.join(work, "fq")
    t_gen = time.perf_counter()
    subprocess.run([synth, "db", "--golden", GOLDEN_B10, "--out", work, "--den", str(den)], check=True,
                   stdout=subprocess.DEVNULL)
    for i in range(n_samples):
        subprocess.run([synth, "reads", "--golden", GOLDEN_B10, "--out", fq, "--sample", "s%03d" % i,
                        "--pairs", str(pairs), "--first-pair", str(i * pairs), "--den", str(den)], check=True,
                       stdout=subprocess.DEVNULL)
    t_gen = time.perf_counter() - t_gen
    t_start = time.perf_counter()
    proc = subprocess.Popen([nk10, fq + "/"], cwd=work, stdout=subprocess.PIPE, text=True)
    stamps = []
    for line in proc.stdout:
        stamps.append((time.perf_counter(), line.rstrip("\n")))
    proc.wait()
    shutil.rmtree(work, ignore_errors=True)
    if proc.returncode != 0:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/nk10 exited %d" % proc.returncode}))
        return
    steps = []
    n_probes = None
    load_s = None
    i = 0
    while i < len(stamps):
        t, text = stamps[i]
        if text.endswith(" kmers loaded"):
            n_probes, load_s = int(text.split()[0]), t - t_start
        if text.startswith("s") and len(text) == 4 and text[1:].isdigit():
            # name line, "<n> reads loaded", "<n> reads loaded"
            t_end, last = stamps[i + 2]
            steps.append((t_end - t, int(last.split()[0])))
            i += 3
        else:
            i += 1
    timed = steps[args.warmup:] if len(steps) > args.warmup else steps
    sec = statistics.mean(s for s, _ in timed)
    value = pairs / sec
    # the arm's own config (the contract: "on your arm's config ... each step a bounded sample of that
    # workload"); what was actually run per step is spelled out in `sample` and cpu_baseline.sample.  A
    # probe DB that is NOT the benchmark's (den != 1) changes the workload and therefore the config.
    cfg = workload_config(args, n_probes, den=den)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u64 keys / int32 counts", "data": "synthetic",
           "config": cfg,
           "sample": {"pairs_per_step": pairs, "db_probes": n_probes, "db_den": den, "steps_in_one_process": n_samples,
                      "db_load_s": load_s, "input_generation_s": t_gen},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "reference",
                            "sample": "unmodified nk10 (g++ -O3), 1 thread (it has no threading); probe DB = refkey counts / %d "
                                      "(%d lines of text)%s in its 2^30-cell table, %d pairs per step (the GPU arm: %d), "
                                      "%d steps in one process; DB load %.0f s and input generation %.0f s are outside the steps"
                                      % (den, n_probes, note, pairs, args.pairs, n_samples, load_s or 0, t_gen)},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def run_reference_port(args):
    """The in-repo CPU port (oracle/kid_oracle.c), one thread: when oracle/_ref/nk10 is absent, and for
    the configs the reference binary cannot run from this command line (mito, x10)."""
    import numpy as np
    from oracle import kor
    den = args.ref_den if args.config != "x10" else max(args.ref_den, 10)  # 1.09 G keys: 35 GB of host RAM
    parent, wl = make_workload(args, den=den)
    keys, taxa = wl.db_host()
    odb = kor.OracleDB(wl.n_taxa)
    odb.set_parents(parent)
    odb.add_keys(keys, taxa)
    n_samples = args.warmup + args.steps
    pairs = max(2000, int(args.ref_seconds * 30000 * 150 / args.read_len / max(1, n_samples)))
    osamp = kor.OracleSample(odb)
    secs = []
    for i in range(n_samples):
        seq, qual = wl.reads_host(2 * i * pairs, 2 * pairs)
        off = wl.offsets(2 * pairs)
        osamp.reset()
        t0 = time.perf_counter()
        osamp.classify(seq, qual, off)
        secs.append(time.perf_counter() - t0)
    sec = statistics.mean(secs[args.warmup:])
    value = pairs / sec
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u64 keys / int32 counts", "data": "synthetic",
           "config": workload_config(args, wl.n_probes, den=den),
           "sample": {"pairs_per_step": pairs, "db_probes": int(keys.size), "db_den": den},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                            "sample": "oracle/kid_oracle.c, probe DB = refkey counts x %d / %d (%d keys), %d pairs per step (the GPU arm: %d)"
                                      % (args.cfg["num"], den, keys.size, pairs, args.pairs)},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
