"""GPU: the data-parallel sharded engine (own queries per rank, DB row-sharded) with both exchanges -- the peer-memory push/wait
kernels over CUDA IPC and NCCL -- equals the float64 oracle over the full DB on every rank.  With >= 2 devices: one rank per
GPU over NCCL.  On a single-GPU box the peer-memory exchange still runs: two PROCESSES share device 0 (legacy CUDA IPC works
between processes on one device; the host-side handshake goes over gloo), so the push / wait kernels, the device-resident
epochs and the captured graphs are exercised wherever the GPU tests run; only the NCCL variant needs two devices."""
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


def _worker(rank, world, port, exchange, q_out, one_gpu=False):
    import torch.distributed as dist

    import oracle
    from text2pos_cvpr2022_b200 import default_args, synthetic as syn
    from text2pos_cvpr2022_b200.cell_retrieval import CellRetrievalNetwork
    from text2pos_cvpr2022_b200.retrieval import shard_bounds
    from text2pos_cvpr2022_b200.serving import OnlineRetrievalEngine, ShardedOnlineRetrievalEngine

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", 0 if one_gpu else rank)
    torch.cuda.set_device(dev)
    if one_gpu:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
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
    one_gpu = torch.cuda.device_count() < 2
    if one_gpu and exchange == "nccl":
        pytest.skip("NCCL needs one device per rank")
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, exchange, q_out, one_gpu)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q_out.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _pipeline_worker(rank, world, port, q_out, one_gpu):
    import sys

    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import pipeline_common as pc
    from text2pos_cvpr2022_b200 import pipeline_eval as pe
    from text2pos_cvpr2022_b200.cell_store import CellStore

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", 0 if one_gpu else rank)
    torch.cuda.set_device(dev)
    if one_gpu:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        ds, loader = pc.scene(11, n_cells=15, n_poses=7)  # odd sizes: unequal shards, a padded last query batch
        args = pc.pipeline_args()
        coarse, _ = pc.coarse_state_dict()
        fine, _ = pc.fine_state_dict()
        coarse, fine = coarse.eval().to(dev), fine.to(dev)
        c_acc, a_mean, a_off, a_conf, info = pe.run_pipeline_distributed(coarse, fine, ds, args, query_batch=3, return_details=True)
        # the same evaluation in ONE process (whole DB, all queries)
        from text2pos_cvpr2022_b200.coarse_eval import eval_epoch_store

        retr, c_ref = pe.run_coarse(coarse, loader, args, eval_epoch_fn=eval_epoch_store)  # DB side on the device data path, seed 0
        store = CellStore.from_cells(ds.all_cells, args.pad_size, lambda cell: pe.seeded_padding_factory(0, cell.id)).to(dev)
        ref = pe.run_fine_cached(fine, retr, loader, args, cache=pe.FineCellCache.from_store(fine, store), return_details=True)
        flat = lambda a: [[float(a[k][t]) for t in sorted(a[k])] for k in sorted(a)]
        q_lo, q_hi = info["query_range"]
        ok = all(list(info["retrievals"][q]) == list(retr[q]) for q in range(q_lo, q_hi))
        ok = ok and flat(c_acc) == flat(c_ref) and flat(a_mean) == flat(ref[0]) and flat(a_off) == flat(ref[1]) and flat(a_conf) == flat(ref[2])
        ok = ok and np.array_equal(info["details"]["matches"], ref[3]["matches"][q_lo:q_hi]) and (ref[3]["matches"] >= 0).sum() > 50
        q_out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_pipeline_config5_distributed_equals_single_process():
    """BASELINE config 5 on 2 ranks (sharded DB build, data-parallel coarse retrieval with the two all-gathers, all-gathered fine
    cache, replica fine stage, all-reduced hit counts) == the single-process pipeline: same retrievals, matches, accuracy dicts."""
    one_gpu = torch.cuda.device_count() < 2
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipeline_worker, args=(r, world, port, q_out, one_gpu)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q_out.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
