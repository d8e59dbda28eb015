#!/usr/bin/env python
"""Diagnosis: the distributed pipeline test body with every comparison printed (2 processes on one GPU over gloo)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.multiprocessing as mp


def worker(rank, world, port):
    import torch.distributed as dist

    import pipeline_common as pc
    from text2pos_cvpr2022_b200 import pipeline_eval as pe
    from text2pos_cvpr2022_b200.cell_store import CellStore

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ds, loader = pc.scene(11, n_cells=15, n_poses=7)
        args = pc.pipeline_args()
        coarse, _ = pc.coarse_state_dict()
        fine, _ = pc.fine_state_dict()
        coarse, fine = coarse.eval().to(dev), fine.to(dev)
        c_acc, a_mean, a_off, a_conf, info = pe.run_pipeline_distributed(coarse, fine, ds, args, query_batch=3, return_details=True)
        from text2pos_cvpr2022_b200.coarse_eval import eval_epoch_store

        retr, c_ref = pe.run_coarse(coarse, loader, args, eval_epoch_fn=eval_epoch_store)  # DB side on the device data path, seed 0
        store = CellStore.from_cells(ds.all_cells, args.pad_size, lambda cell: pe.seeded_padding_factory(0, cell.id)).to(dev)
        ref = pe.run_fine_cached(fine, retr, loader, args, cache=pe.FineCellCache.from_store(fine, store), return_details=True)
        flat = lambda a: [[float(a[k][t]) for t in sorted(a[k])] for k in sorted(a)]
        q_lo, q_hi = info["query_range"]
        print(rank, "range", q_lo, q_hi, "retr", [list(info["retrievals"][q]) == list(retr[q]) for q in range(q_lo, q_hi)], flush=True)
        for q in range(q_lo, q_hi):
            if list(info["retrievals"][q]) != list(retr[q]):
                print(rank, q, list(info["retrievals"][q]), list(retr[q]), flush=True)
        print(rank, "coarse", flat(c_acc), flat(c_ref), flush=True)
        print(rank, "mean", flat(a_mean) == flat(ref[0]), "off", flat(a_off) == flat(ref[1]), "conf", flat(a_conf) == flat(ref[2]), flush=True)
        print(rank, "matches", np.array_equal(info["details"]["matches"], ref[3]["matches"][q_lo:q_hi]), (ref[3]["matches"] >= 0).sum(), flush=True)
        full = build = None
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    ctx = mp.get_context("spawn")
    ps = [ctx.Process(target=worker, args=(r, 2, 29577)) for r in range(2)]
    for p in ps:
        p.start()
    for p in ps:
        p.join()
