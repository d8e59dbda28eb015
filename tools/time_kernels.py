#!/usr/bin/env python
"""Kernel-level timings of the online path on the GPU box (CUDA events, L2-cold DB copies), one line per variant.

    python tools/time_kernels.py [--iters 200]

Prints: text encoder per path (1 = shared-memory cluster kernel, 2 = register kernel, 3 = tensor-core kernel) with the
max abs error of each against path 2, and the retrieval top-k at the BASELINE DB sizes.
"""
import argparse
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def time_cuda(fn, iters, warmup=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i) if fn.__code__.co_argcount else fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    args = ap.parse_args()
    import bench
    from text2pos_cvpr2022_b200 import _lib, synthetic as syn
    from text2pos_cvpr2022_b200.modules import tokenize
    from text2pos_cvpr2022_b200.serving import OnlineRetrievalEngine

    dev = torch.device("cuda", 0)
    model = bench.build_model().to(dev)
    kw = model.language_encoder.known_words
    texts = syn.synth_queries(1000, 64)
    t, l = tokenize(texts, kw)
    d_tok = torch.from_numpy(t.astype(np.int32)).to(dev)
    d_len = torch.from_numpy(l.astype(np.int32)).to(dev)
    base = syn.synth_db_embeddings(100, 10000, 256).to(dev)
    eng = OnlineRetrievalEngine(model, base, k=10, max_batch=64, max_tokens=t.shape[1])
    outs = {}
    for path in (2, 3, 1):
        eng.lstm_desc = copy.copy(eng.lstm_desc)
        eng.lstm_desc.path = path
        try:
            us = time_cuda(lambda: eng.enqueue_encode(d_tok, d_len), args.iters)
        except Exception as ex:  # noqa: BLE001
            print(f"lstm path {path}: {ex}")
            continue
        outs[path] = eng.q.clone()
        err = float((outs[path] - outs[2]).abs().max()) if 2 in outs else float("nan")
        print(f"lstm path {path}: {us:8.1f} us / batch of 64 x {t.shape[1]} tokens   max|diff vs path 2| = {err:.2e}", flush=True)
    eng.lstm_desc.path = 0
    eng.enqueue_encode(d_tok, d_len)
    # top-k at the per-rank shapes of the data-parallel sharded engine: R*64 queries against a 12,500-row shard
    from text2pos_cvpr2022_b200.retrieval import retrieve_topk, db_row_norm2_max
    db = syn.synth_db_embeddings(100, 12500, 256).to(dev)
    copies = [db.clone() for _ in range(32)]
    nm = db_row_norm2_max(db)
    ws = _lib.Workspace()
    for B in (64, 128, 256, 512):
        q = syn.synth_query_embeddings(3, B, 256).to(dev)
        us = time_cuda(lambda i=0: retrieve_topk(q, copies[i % 32], 10, 0, ws, nm), args.iters)
        print(f"topk B={B:4d} N=12500: {us:8.1f} us (incl. 2 torch.empty per call)", flush=True)
    for n in (10000, 12500, 100000):
        db = syn.synth_db_embeddings(100, n, 256).to(dev)
        ncopies = max(2, int(400e6 // (n * 1024)) + 1)
        copies = [db.clone() for _ in range(ncopies)]
        eng.set_db(db)
        us = time_cuda(lambda i=0: eng.enqueue_topk(copies[i % ncopies]), args.iters)
        gbs = (n * 1024 + 64 * 1024 + 64 * 160) / us / 1e3
        print(f"topk N={n:6d}: {us:8.1f} us  ({gbs:7.1f} GB/s algorithmic, {ncopies} rotating copies)", flush=True)


if __name__ == "__main__":
    main()
