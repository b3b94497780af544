"""Run under torchrun (one rank per GPU): each rank classifies its own shard of one sample, the
sample-end reduction (kmer_id_b200/multi_gpu.py over NCCL) must reproduce the oracle's gcount/ucount
for the WHOLE sample on every rank.  Exit code 0 = parity."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import kmer_id_b200 as kid
    from kmer_id_b200 import multi_gpu
    from oracle import kor
    from tools import synthlib

    parent, prefix = synthlib.load_taxonomy(os.path.join(ROOT, "tests", "golden", "b10"), 1, 200)
    wl = synthlib.Workload(parent, prefix)
    keys, taxa = wl.db_host()
    n_per = 30000
    stream = torch.cuda.current_stream().cuda_stream
    for layout, peer in ((0, False), (kid.KID_DB_LAYOUT_KEYHASH, False), (0, True)):
        db = kid.Database(keys, taxa, parent, device=local, flags=layout)
        s = kid.Sample(db)
        engine, transport = multi_gpu.make_engine(s, stream, prefer_peer=peer)
        if peer and rank == 0:
            print("transport for the peer case:", transport)
        seq, qual = wl.reads_host(rank * n_per, n_per)
        # every rank also sees the same first 2000 reads: their k-mers must count once in ucount
        cs, cq = wl.reads_host(10_000_000, 2000)
        seq = np.concatenate([seq[:n_per * 150], cs])
        qual = np.concatenate([qual[:n_per * 150], cq])
        n = n_per + 2000
        off = wl.offsets(n)
        s.begin(stream)
        out = s.classify(seq, qual, off)
        g, u = multi_gpu.finish(engine)
        # oracle over the union of all shards (each rank computes it; small)
        odb = kor.OracleDB(wl.n_taxa)
        odb.set_parents(parent)
        odb.add_keys(keys, taxa)
        os_ = kor.OracleSample(odb)
        for r in range(world):
            sq, ql = wl.reads_host(r * n_per, n_per)
            sq = np.concatenate([sq[:n_per * 150], cs])
            ql = np.concatenate([ql[:n_per * 150], cq])
            fin, _ = os_.classify(sq, ql, off)
            if r == rank:
                assert np.array_equal(fin, out), "per-read taxa differ from the oracle"
        assert np.array_equal(g, os_.gcount), f"rank {rank}: gcount differs after all-reduce"
        assert np.array_equal(u, os_.ucount), f"rank {rank}: ucount differs after OR-reduce"
        assert u.sum() > 1000
        del engine, s, db
    dist.barrier()
    if rank == 0:
        print(f"multi-GPU parity ok on {world} ranks (both layouts)")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
