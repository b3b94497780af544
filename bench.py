#!/usr/bin/env python
"""bench.py - the nk10 read-classification hot path on B200, one process per GPU.

    python bench.py --gpus N --steps K --warmup W              (ours; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU path)

Workload (BASELINE.json configs[2], the largest single-GPU configuration): synthetic bact10-scale
probe database (108 585 519 probes = sum of the shipped refkey10.txt counts, b10 taxonomy) and
10 M synthetic 150-bp read pairs per GPU (tools/synth/kid_synth.h).  A *step* is one whole sample:
kid_sample_begin -> classify every read of the batch -> sample end (ucount histogram, counts to
the host; for N > 1 the cross-rank gcount sum / seen-bitmap OR).  A pair is two independently
classified reads.

  value  pairs/s with the batch already resident in HBM (CUDA events on the launching stream)
  e2e    the same through kid_classify_host from pinned HOST buffers (H2D + D2H inside the region)
  roofline.achieved = lookups/launch x 32 B / mean classify-kernel time  (32 B = one DRAM sector per
           lookup); peak = MEASURED_PEAKS.json hbm_gbs
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
METRIC = "paired_reads_per_sec_classified"
UNIT = "pairs/s"
READ_LEN = 150
SECTOR_BYTES = 32


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=10_000_000, help="read pairs per GPU per step")
    ap.add_argument("--db-den", type=int, default=1, help="probe DB = refkey counts / den")
    ap.add_argument("--cpu-baseline-reads", type=int, default=400_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-files-e2e", action="store_true",
                    help="skip the gz-files-on-disk -> _result.txt run of the nk10 drop-in")
    ap.add_argument("--layout", default="M", choices=["M", "K"], help="table layout (M = minimizer, default)")
    ap.add_argument("--log2-sectors", type=int, default=0, help="table size override (0 = library default)")
    ap.add_argument("--ref-seconds", type=float, default=80.0,
                    help="CPU seconds the reference arm may spend classifying (all steps together)")
    return ap.parse_args()


def workload_config(args, n_probes):
    return {
        "workload": "BASELINE.json configs[2]: bact10-scale synthetic probe DB + 10M synthetic 150bp pairs per GPU",
        "db_probes": int(n_probes),
        "db": "b10 taxonomy, refkey10 probe counts / %d, random canonical 30-mers (seed 10)" % args.db_den,
        "pairs_per_gpu_per_step": args.pairs,
        "read_len": READ_LEN,
        "reads": "70% stitched from lineage probes, 0.5% subs, 0.1% N, 20% low-quality tails (seed 21)",
        "sharding": "reads sharded across ranks, table replicated",
        "l2": "inputs_larger_than_l2 (%.1f GB of reads per step, GB-scale probe table: see table.bytes)"
              % (args.pairs * 2 * READ_LEN * 2 / 1e9),
    }


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
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import kmer_id_b200 as kid  # raises if the CUDA library is not built: no fallback
    from kmer_id_b200 import multi_gpu
    from tools import synthlib

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

    parent, prefix = synthlib.load_taxonomy(GOLDEN_B10, 1, args.db_den)
    wl = synthlib.Workload(parent, prefix, read_len=READ_LEN)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- database: generated on the device, built by the library's own kernels
    launches0 = kid.kernel_launches()
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

    # ---- this rank's reads, resident in HBM
    n_reads = 2 * args.pairs
    nbytes = n_reads * READ_LEN
    dseq = torch.empty(nbytes + 64, dtype=torch.uint8, device=dev)
    dqual = torch.empty(nbytes + 64, dtype=torch.uint8, device=dev)
    wl.reads_device(local, rank * n_reads, n_reads, dseq, dqual, stream=stream)
    doff = torch.arange(n_reads + 1, dtype=torch.int64, device=dev) * READ_LEN
    dout = torch.empty(n_reads, dtype=torch.int32, device=dev)
    sample = kid.Sample(db)
    engine, transport = multi_gpu.make_engine(sample, stream, prefer_peer=os.environ.get("KID_PEER", "1") != "0")
    torch.cuda.synchronize()

    k_ev = []

    def step_device(timed: bool):
        sample.begin(stream)
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        sample.classify_device(dseq, dqual, doff, n_reads, dout, None, stream)
        if timed:
            b.record()
            k_ev.append((a, b))
        return multi_gpu.finish(engine)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        gcount, ucount = step_device(False)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.3)
    l_before = kid.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tc0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        gcount, ucount = step_device(True)
    e1.record()
    barrier()
    tc1 = time.perf_counter()
    gpu_launches = kid.kernel_launches() - l_before
    ms = e0.elapsed_time(e1)
    counters = sample.counters(stream)  # of the last step
    kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in k_ev)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * args.pairs / (ms_per_step / 1e3)
    if world == 1:
        assert int(gcount.sum()) == counters["reads"], "gcount does not add up to the reads classified"

    # ---- end to end through the host-buffer entry point
    e2e = None
    if not args.no_e2e:
        hseq = torch.empty(nbytes + 64, dtype=torch.uint8, pin_memory=True)
        hqual = torch.empty(nbytes + 64, dtype=torch.uint8, pin_memory=True)
        hseq.copy_(dseq)
        hqual.copy_(dqual)
        hoff = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(READ_LEN))
        hout = torch.empty(n_reads, dtype=torch.int32, pin_memory=True)
        torch.cuda.synchronize()

        def step_e2e():
            sample.begin(stream)
            sample.classify_host(hseq, hqual, hoff, n_reads, hout, None)
            return multi_gpu.finish(engine)

        for _ in range(max(1, min(args.warmup, 2))):
            g2, u2 = step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            g2, u2 = step_e2e()
        barrier()
        dt_e2e = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt_e2e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_e2e = float(t.item())
        h2d, d2h = sample.transfer_bytes()
        assert np.array_equal(g2, gcount) and np.array_equal(u2, ucount), "host and device paths disagree"
        assert np.array_equal(hout.numpy(), dout.cpu().numpy())
        e2e = {"value": world * args.pairs * args.steps / dt_e2e, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h + 8 * db.n_taxa),
               "ms_per_step": dt_e2e / args.steps * 1e3}
    clk = clocks.stop(tc0, None)

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
    roofline = {"bound": "hbm", "kernel": "kid_classify2_kernel" if args.layout == "M" else "kid_classify_kernel", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "traffic": (tr["dram_bytes_per_lookup"] * lookups if tr else None),
                "algorithmic_bytes_per_lookup": SECTOR_BYTES, "lookups_per_launch": lookups,
                "kernel_ms": kernel_ms, "lookups_per_s": lookups / (kernel_ms / 1e3),
                "random_sector_gather_ceiling_gsectors_s": ceiling, "gather_ceiling_source": ceiling_src}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 keys / int32 counts", "data": "synthetic",
        "config": workload_config(args, wl.n_probes), "clocks": clk, "e2e": e2e,
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
    }

    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_port(args, wl, dk, dt, parent, dout.cpu().numpy())
    if world == 1 and not args.no_files_e2e:
        out["files_e2e"] = files_e2e(args)
    del dk, dt
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def files_e2e(args):
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
        for i in range(3):
            subprocess.run([synth, "reads", "--golden", GOLDEN_B10, "--out", fq, "--sample", "s%d" % i, "--pairs",
                            str(pairs), "--first-pair", str(i * pairs), "--den", str(args.db_den)], check=True,
                           stdout=subprocess.DEVNULL)
        t0 = time.perf_counter()
        r = subprocess.run([nk10, fq + "/"], cwd=work, capture_output=True, text=True,
                           env=dict(os.environ, KID_STATS="1", KID_NO_CACHE="1"))
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {"error": "nk10 exited %d: %s" % (r.returncode, r.stderr[-300:])}
        per_sample = [float(x) for x in re.findall(r"\[nk10\] s\d+: \d+ reads, \d+ lookups, \d+ hits in ([0-9.]+) s", r.stderr)]
        m = re.search(r"\[nk10\] parse db ([0-9.]+) s, build table ([0-9.]+) s, total ([0-9.]+) s", r.stderr)
        best = min(per_sample)
        return {"value": pairs / best, "unit": UNIT,
                "what": "kmer_id_b200/bin/nk10: gz FASTQ on disk -> _result.txt/_reads.txt, fastest of 3 samples "
                        "(R1 and R2 inflated on all host cores, parsed, classified); probe DB parsed from gz text",
                "pairs_per_sample": pairs, "sample_s": per_sample, "db_parse_s": float(m.group(1)),
                "table_build_s": float(m.group(2)), "process_total_s": float(m.group(3)), "wall_s": wall,
                "whole_run_pairs_per_s": 3 * pairs / wall, "host_cores": os.cpu_count()}
    finally:
        shutil.rmtree(work, ignore_errors=True)


def cpu_baseline_port(args, wl, dk, dt, parent, gpu_taxa):
    """The CPU oracle (kind = "port"), one thread, on the first cpu_baseline_reads reads of the same
    workload against the FULL probe table (so its cache behaviour is the real one).  Its per-read
    taxa double as a parity check of what the timed GPU steps computed for those reads."""
    import numpy as np
    from oracle import kor
    n = min(args.cpu_baseline_reads, 2 * args.pairs)
    keys = dk.cpu().numpy().view(np.uint64)
    taxa = dt.cpu().numpy().view(np.uint32)
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
    return {"value": (n / 2) / dt_s, "unit": UNIT, "cores": 1, "kind": "port", "parity_checked_reads": int(n),
            "sample": "first %d reads of rank 0's batch, full %d-probe table, oracle/kid_oracle.c" % (n, keys.size),
            "lookups_per_s": osamp.lookups / dt_s, "seconds": dt_s}


# ------------------------------------------------------------------------------------ reference
def run_reference(args):
    """The reference's own CPU implementation: oracle/_ref/nk10 (unmodified newkmer_10nx.cpp,
    single-threaded - it has no threading) on a bounded sample of the workload: a 1/100-scale probe
    DB in its own text format and (warmup+steps) samples of P pairs each, one nk10 process; a step is
    one sample, timed from its name line to its second "reads loaded" line on stdout."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import tempfile
    nk10 = os.path.join(ROOT, "oracle", "_ref", "nk10")
    synth = os.path.join(ROOT, "tools", "kid_synth")
    n_samples = args.warmup + args.steps
    if not (os.path.exists(nk10) and os.path.exists(synth)):
        return run_reference_port(args)
    pairs = max(2000, int(args.ref_seconds * 16000 / max(1, n_samples)))  # ~32 k reads/s single thread
    den = 100
    work = tempfile.mkdtemp(prefix="kid_ref_")
    fq = os.path.join(work, "fq")
    subprocess.run([synth, "db", "--golden", GOLDEN_B10, "--out", work, "--den", str(den)], check=True,
                   stdout=subprocess.DEVNULL)
    for i in range(n_samples):
        subprocess.run([synth, "reads", "--golden", GOLDEN_B10, "--out", fq, "--sample", "s%03d" % i,
                        "--pairs", str(pairs), "--first-pair", str(i * pairs), "--den", str(den)], check=True,
                       stdout=subprocess.DEVNULL)
    proc = subprocess.Popen([nk10, fq + "/"], cwd=work, stdout=subprocess.PIPE, text=True)
    stamps = []
    for line in proc.stdout:
        stamps.append((time.perf_counter(), line.rstrip("\n")))
    proc.wait()
    if proc.returncode != 0:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/nk10 exited %d" % proc.returncode}))
        return
    steps = []
    i = 0
    while i < len(stamps):
        t, text = stamps[i]
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
    import shutil
    shutil.rmtree(work, ignore_errors=True)
    cfg = workload_config(args, 108_585_519)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u64 keys / int32 counts", "data": "synthetic",
           "config": cfg,
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "reference",
                            "sample": "unmodified nk10 (g++ -O3), 1 thread; probe DB sampled 1/%d (%s lines of text) in "
                                      "its 2^30-cell table, %d pairs per step, %d steps in one process"
                                      % (den, "1083547", pairs, n_samples)},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def run_reference_port(args):
    """Fallback when oracle/_ref/nk10 is absent: the in-repo CPU port, one thread."""
    import numpy as np
    from oracle import kor
    from tools import synthlib
    parent, prefix = synthlib.load_taxonomy(GOLDEN_B10, 1, 100)
    wl = synthlib.Workload(parent, prefix, read_len=READ_LEN)
    keys, taxa = wl.db_host()
    odb = kor.OracleDB(wl.n_taxa)
    odb.set_parents(parent)
    odb.add_keys(keys, taxa)
    n_samples = args.warmup + args.steps
    pairs = max(2000, int(args.ref_seconds * 30000 / max(1, n_samples)))
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
           "config": workload_config(args, 108_585_519),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                            "sample": "oracle/kid_oracle.c, 1/100 probe DB, %d pairs per step" % pairs},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
