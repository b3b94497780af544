"""world_size-2 (and 4) gloo test of the sample-end reduction (kmer_id_b200/multi_gpu.py) on CPU:
the orchestration that runs under NCCL on the GPUs, with numpy stand-ins for the two kernels.
Checks the property SURVEY.md fact 3 demands: ucount of the sharded run == ucount of the OR of all
shards' seen flags (NOT the sum of per-shard ucounts)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H  # noqa: F401  (sys.path)


class CpuEngine:
    def __init__(self, gcount, seen_bits_words, slot_taxon):
        self.n_taxa = gcount.size
        self.gcount = torch.from_numpy(gcount.copy())
        self.seen = torch.from_numpy(seen_bits_words.copy().view(np.int32))
        self.n_words = seen_bits_words.size
        self.slot_taxon = slot_taxon

    def new_partial(self):
        return torch.zeros(self.n_taxa, dtype=torch.int32)

    def new_recv(self):
        return torch.empty(self.n_words, dtype=torch.int32)

    def or_into_own(self, recv, word0, n_words, world):
        acc = np.zeros(n_words, np.uint32)
        r = recv.numpy().view(np.uint32)
        for k in range(world):
            acc |= r[k * n_words:(k + 1) * n_words]
        self.seen.numpy().view(np.uint32)[word0:word0 + n_words] = acc

    def ucount_range(self, word0, n_words, partial):
        w = self.seen.numpy().view(np.uint32)[word0:word0 + n_words]
        bits = np.unpackbits(w.view(np.uint8), bitorder="little")
        slots = np.flatnonzero(bits) + 32 * word0
        np.add.at(partial.numpy(), self.slot_taxon[slots], 1)


def _make_case(world, seed=5):
    rng = np.random.default_rng(seed)
    n_words, n_taxa = 4096, 50
    slot_taxon = rng.integers(2, n_taxa, size=32 * n_words).astype(np.int64)
    shards = []
    for r in range(world):
        seen = np.zeros(n_words, np.uint32)
        hit = rng.choice(32 * n_words, size=3000, replace=False)
        common = np.arange(0, 32 * n_words, 97)  # slots every rank sees: must count once
        for s in np.concatenate([hit, common]):
            seen[s >> 5] |= np.uint32(1 << (s & 31))
        g = rng.integers(0, 1000, size=n_taxa).astype(np.int32)
        shards.append((g, seen))
    return slot_taxon, shards, n_taxa


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kmer_id_b200 import multi_gpu
    slot_taxon, shards, n_taxa = _make_case(world)
    g, seen = shards[rank]
    eng = CpuEngine(g, seen, slot_taxon)
    gc, uc = multi_gpu.sample_end(eng)
    q.put((rank, gc, uc))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sample_end_matches_single_rank_union(world):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    slot_taxon, shards, n_taxa = _make_case(world)
    want_g = sum(g.astype(np.int64) for g, _ in shards).astype(np.int32)
    union = np.zeros_like(shards[0][1])
    for _, seen in shards:
        union |= seen
    bits = np.unpackbits(union.view(np.uint8), bitorder="little")
    want_u = np.bincount(slot_taxon[np.flatnonzero(bits)], minlength=n_taxa).astype(np.int32)
    naive_sum = sum(np.bincount(slot_taxon[np.flatnonzero(np.unpackbits(s.view(np.uint8), bitorder="little"))],
                                minlength=n_taxa) for _, s in shards)
    assert (naive_sum != want_u).any(), "fixture must make the naive sum wrong"
    for rank, gc, uc in outs:
        assert np.array_equal(gc, want_g)
        assert np.array_equal(uc, want_u)


def test_single_rank_path():
    from kmer_id_b200 import multi_gpu
    slot_taxon, shards, n_taxa = _make_case(1)
    g, seen = shards[0]
    gc, uc = multi_gpu.sample_end(CpuEngine(g, seen, slot_taxon))
    bits = np.unpackbits(seen.view(np.uint8), bitorder="little")
    assert np.array_equal(gc, g)
    assert np.array_equal(uc, np.bincount(slot_taxon[np.flatnonzero(bits)], minlength=n_taxa))
