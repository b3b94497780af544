"""Sample-end reduction across ranks (one process per GPU, torch.distributed for the plumbing).

The reference has one global ``gcount[]``/``ucount[]``/``kmer_seen`` (newkmer_10nx.cpp:61-64).  When
reads are sharded over ranks:

* ``gcount`` is additive              -> one sum all-reduce of int32[n_taxa];
* ``ucount`` is NOT additive (a k-mer seen on two ranks must count once).  It is
  the per-taxon histogram of the OR of every rank's per-slot seen bitmap.  The bitmap is cut into
  world_size equal word ranges; an all-to-all hands rank r everybody's range r, rank r ORs them
  into its own bitmap and histograms that range only.  The ranges are disjoint, so the partial
  histograms ARE additive and a second small sum all-reduce finishes ucount.

Two transports for the bitmap exchange: NCCL all-to-all + kid_seen_or_kernel (sample_end), or - when
the ranks can map each other's memory - symmetric memory and ONE fused OR+histogram kernel that
reads the peers' bitmaps over NVLink in place (sample_end_peer).

The orchestration is written against a tiny "engine" interface so that the same code runs on CPU
tensors under gloo in tests/test_multi_rank_cpu.py and on the CUDA kernels under NCCL.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


class _DevArray:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False),
                                         "version": 2}


def wrap_device(ptr: int, n: int, typestr: str, device) -> torch.Tensor:
    return torch.as_tensor(_DevArray(ptr, n, typestr), device=device)


class CudaEngine:
    """Adapter from kmer_id_b200.Sample to the engine interface used by sample_end()."""

    def __init__(self, sample, stream: int = 0):
        self.sample = sample
        self.stream = stream
        self.device = torch.device("cuda", sample.db.device)
        self.n_taxa = sample.db.n_taxa
        gp = sample.gcount_device()
        sp, self.n_words = sample.seen_device()
        self._seen_ptr = sp
        self.gcount = wrap_device(gp, self.n_taxa, "<i4", self.device)
        self.seen = wrap_device(sp, self.n_words, "<i4", self.device)

    def new_partial(self) -> torch.Tensor:
        return torch.zeros(self.n_taxa, dtype=torch.int32, device=self.device)

    def new_recv(self) -> torch.Tensor:
        return torch.empty(self.n_words, dtype=torch.int32, device=self.device)

    def or_into_own(self, recv: torch.Tensor, word0: int, n_words: int, world: int):
        srcs = [recv.data_ptr() + 4 * k * n_words for k in range(world)]
        self.sample.seen_or(self._seen_ptr + 4 * word0, srcs, 0, n_words, self.stream)

    def ucount_range(self, word0: int, n_words: int, partial: torch.Tensor):
        self.sample.ucount_range(self._seen_ptr, word0, n_words, partial, self.stream)


class PeerEngine(CudaEngine):
    """Seen bitmap in symmetric memory: every rank maps every peer's bitmap over NVLink, so the
    sample-end OR-reduction and the per-taxon histogram are ONE kernel reading peer memory in place
    (kid_ucount_or_range_device) - no all-to-all, no staging copy, nothing written back."""

    def __init__(self, sample, stream: int = 0, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        super().__init__(sample, stream)
        group = group or dist.group.WORLD
        self.buf = symm_mem.empty(self.n_words, dtype=torch.int32, device=self.device)
        self.hdl = symm_mem.rendezvous(self.buf, group.group_name)
        self.peer_ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        assert len(self.peer_ptrs) == dist.get_world_size(group)
        sample.use_seen_buffer(self.buf.data_ptr(), self.n_words)
        self.seen = self.buf
        self._seen_ptr = self.buf.data_ptr()
        sample.begin(stream)  # the new buffer starts cleared


def sample_end_peer(engine: PeerEngine, group=None):
    """(gcount, ucount) through peer memory; identical on every rank."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    assert engine.n_words % (4 * world) == 0
    sw = engine.n_words // world
    partial = engine.new_partial()
    engine.hdl.barrier(channel=0)  # every rank's classify kernels have finished writing seen bits
    engine.sample.ucount_or_range(engine.peer_ptrs, rank * sw, sw, partial, engine.stream)
    engine.hdl.barrier(channel=1)  # peers are done reading this bitmap before the next begin() clears it
    dist.all_reduce(engine.gcount, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return engine.gcount.cpu().numpy().copy(), partial.cpu().numpy()


def check_replicas_agree(sample, group=None):
    """Slot i must mean the same key on every rank: the bitmaps are OR-ed by slot index.  kid_db_build
    sizes the table from the key count alone and places keys deterministically, so replicas built from
    the same probe list agree - unless a device was short of memory and took a smaller table.  Raise
    instead of exchanging misaligned bitmaps."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    st = sample.db.stats()
    mine = torch.tensor([st["n_sectors"], st["table_bytes"], st["n_distinct"], st["n_displaced"],
                         sample.seen_device()[1], sample.db.n_taxa], dtype=torch.int64,
                        device=torch.device("cuda", sample.db.device))
    every = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
    dist.all_gather(every, mine, group=group)
    for r, other in enumerate(every):
        if not torch.equal(other, mine):
            raise RuntimeError(f"kmer_id_b200.multi_gpu: rank {r}'s table replica differs from rank "
                               f"{dist.get_rank(group)}'s ({other.tolist()} vs {mine.tolist()}): "
                               "build every replica with the same log2_sectors")


def make_engine(sample, stream: int = 0, group=None, prefer_peer: bool = True):
    """PeerEngine when the ranks can map each other's memory, else the NCCL all-to-all engine."""
    check_replicas_agree(sample, group)
    if prefer_peer and dist.is_initialized() and dist.get_world_size(group) > 1:
        try:
            return PeerEngine(sample, stream, group), "peer"
        except Exception as e:  # no P2P / symmetric memory on this box
            import sys
            print(f"kmer_id_b200.multi_gpu: peer memory unavailable ({type(e).__name__}: {e}); using NCCL all-to-all",
                  file=sys.stderr)
    return CudaEngine(sample, stream), "nccl"


def finish(engine, group=None):
    return sample_end_peer(engine, group) if isinstance(engine, PeerEngine) else sample_end(engine, group)


def sample_end(engine, group=None):
    """Returns (gcount, ucount) as int32 numpy arrays, identical on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    partial = engine.new_partial()
    if world == 1:
        engine.ucount_range(0, engine.n_words, partial)
        return engine.gcount.cpu().numpy().copy(), partial.cpu().numpy()
    assert engine.n_words % (4 * world) == 0, "seen bitmap must split into 4-word multiples"
    sw = engine.n_words // world
    dist.all_reduce(engine.gcount, op=dist.ReduceOp.SUM, group=group)
    recv = engine.new_recv()
    dist.all_to_all_single(recv, engine.seen, group=group)
    engine.or_into_own(recv, rank * sw, sw, world)
    engine.ucount_range(rank * sw, sw, partial)
    dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return engine.gcount.cpu().numpy().copy(), partial.cpu().numpy()
