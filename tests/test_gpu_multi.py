"""GPU, >= 2 devices: the data-parallel sharded engine (own queries per rank, DB row-sharded) with both exchanges --
the peer-memory push/wait kernels over CUDA IPC and NCCL -- equals the float64 oracle over the full DB on every rank.
Skipped on single-GPU boxes."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, exchange, q_out):
    import torch.distributed as dist

    import oracle
    from text2pos_cvpr2022_b200 import default_args, synthetic as syn
    from text2pos_cvpr2022_b200.cell_retrieval import CellRetrievalNetwork
    from text2pos_cvpr2022_b200.retrieval import shard_bounds
    from text2pos_cvpr2022_b200.serving import OnlineRetrievalEngine, ShardedOnlineRetrievalEngine

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        model = CellRetrievalNetwork(syn.KNOWN_CLASSES, syn.COLOR_NAMES, syn.known_words(), default_args(embed_dim=256))
        syn.randomize_module_(model, 5, gain=2.0)
        model = model.eval().to(dev)
        n, B, k = 3001, 8, 10
        db = syn.synth_db_embeddings(7, n, 256)
        lo, hi = shard_bounds(n, world)[rank]
        eng = OnlineRetrievalEngine(model, db[lo:hi].to(dev), k=k, max_batch=B, max_tokens=64, idx_base=lo, depth=2)
        sh = ShardedOnlineRetrievalEngine(eng, exchange=exchange)
        ok = True
        batches = [syn.synth_queries(100 + 10 * rank + i, B) for i in range(5)]
        want = []
        for b in batches:  # synchronous calls; reference = oracle over the FULL DB with this rank's own embeddings
            i, s = sh.query(b)
            q = eng.slots[0].q.cpu().numpy()
            ri, rs = oracle.retrieval.topk(db.numpy(), q, k)
            ok = ok and np.array_equal(i, ri) and np.allclose(s, rs, rtol=1e-12, atol=0)
            want.append(i.copy())
        if exchange == "p2p":
            sh.capture_all("g")
        got = []
        for j, b in enumerate(batches):  # two batches in flight (graph replays on odd steps in p2p mode)
            if len(sh._inflight) == 2:
                got.append(sh.collect()[0].copy())
            sh.submit(b, graph_key="g" if j % 2 else None)
        while sh._inflight:
            got.append(sh.collect()[0].copy())
        ok = ok and all(np.array_equal(a, b) for a, b in zip(got, want))
        sh.close()
        q_out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_sharded_engine_data_parallel(exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, exchange, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q_out.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
