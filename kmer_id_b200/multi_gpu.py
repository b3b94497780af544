"""Sample-end reduction across ranks (one process per GPU, torch.distributed for the plumbing).

The reference has one global ``gcount[]``/``ucount[]``/``kmer_seen`` (newkmer_10nx.cpp:61-64).  When
reads are sharded over ranks:

* ``gcount`` is additive              -> one sum all-reduce of int32[n_taxa];
* ``ucount`` is NOT additive (a k-mer seen on two ranks must count once).  It is
  the per-taxon histogram of the OR of every rank's per-slot seen bitmap.  The bitmap is cut into
  world_size equal word ranges; an all-to-all hands rank r everybody's range r, rank r ORs them
  into its own bitmap and histograms that range only.  The ranges are disjoint, so the partial
  histograms ARE additive and a second small sum all-reduce finishes ucount.

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
