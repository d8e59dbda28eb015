"""CPU, world_size 2 over gloo: the shard / all-gather / merge plumbing of ShardedCellDatabase.

The CUDA kernels cannot run here, so the local top-k and the merge are injected from the oracle; what is tested is
the partitioning (row blocks, idx_base), the single packed all-gather and that the sharded result equals the
unsharded oracle on every rank.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from text2pos_cvpr2022_b200 import synthetic as syn
from text2pos_cvpr2022_b200.retrieval import ShardedCellDatabase, shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_local(q, db, k, base):
    idx, sc = oracle.retrieval.topk(db.numpy(), q.numpy(), min(k, db.shape[0]))
    i = torch.full((q.shape[0], k), -1, dtype=torch.int64)
    s = torch.full((q.shape[0], k), float("-inf"), dtype=torch.float64)
    i[:, : idx.shape[1]] = torch.from_numpy(idx) + base
    s[:, : idx.shape[1]] = torch.from_numpy(sc)
    return i, s


def _oracle_merge(gs, gi, k):
    R = gs.shape[0]
    # -1 (empty) entries carry -inf scores; give them a huge index so that they sort last among equals
    gi2 = torch.where(gi < 0, torch.full_like(gi, 2**62), gi)
    mi, ms = oracle.retrieval.merge_shards([gi2[r].numpy() for r in range(R)], [gs[r].numpy() for r in range(R)], k)
    return torch.from_numpy(mi), torch.from_numpy(ms)


def _worker(rank, world, port, n, q_out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        db = syn.synth_db_embeddings(1, n, 64)
        q = syn.synth_query_embeddings(2, 5, 64)
        lo, hi = shard_bounds(n, world)[rank]
        sdb = ShardedCellDatabase(db[lo:hi], n, local_topk=_oracle_local, merge=_oracle_merge)
        idx, sc = sdb.topk(q, 10)
        ref_i, ref_s = oracle.retrieval.topk(db.numpy(), q.numpy(), 10)
        ok = np.array_equal(idx.numpy(), ref_i) and np.allclose(sc.numpy(), ref_s, rtol=1e-13, atol=0)
        # data-parallel queries: every rank brings its own batch and gets the global top-k of its own rows
        q_mine = syn.synth_query_embeddings(10 + rank, 4, 64)
        idx2, sc2 = sdb.topk_dp(q_mine, 10)
        ref_i2, ref_s2 = oracle.retrieval.topk(db.numpy(), q_mine.numpy(), 10)
        ok = ok and np.array_equal(idx2.numpy(), ref_i2) and np.allclose(sc2.numpy(), ref_s2, rtol=1e-13, atol=0)
        q_out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 13])
def test_sharded_topk_equals_unsharded(n):
    world = 2
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q_out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
