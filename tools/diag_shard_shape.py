"""One-GPU emulation of the PER-GPU kernel work of the sharded serving step at R ranks (diagnostic, never a bench value).

Every step: device tokeniser + text encoder of the own 64 queries, top-k of R*64 queries against a 12,500-row shard, merge of R
lists.  The two NVLink exchanges are replaced by local copies, so the difference between this number and the real N = R step
is what the exchanges, the cross-rank skew and the host cost.   usage: python tools/diag_shard_shape.py [R] [depth] [K]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text2pos_cvpr2022_b200 import _lib  # noqa: E402

if os.environ.get("T2P_DIAG_LIB"):  # an experimental build of the library (trace / knock-out experiments)
    _lib.LIB_PATH = os.environ["T2P_DIAG_LIB"]
import bench  # noqa: E402
from text2pos_cvpr2022_b200 import synthetic as syn  # noqa: E402
from text2pos_cvpr2022_b200.serving import OnlineRetrievalEngine  # noqa: E402


def run(R, depth, K, scan_ctas, lstm_clusters=None, n_rows=12500, only=None, timeline=False, native=False, ncopies=None):
    NC = ncopies or bench.N_DB_COPIES
    import ctypes
    stamp_lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin", "libstamp.so")) if timeline else None
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    model = bench.build_model().to(dev)
    B, k, D = bench.B_QUERIES, bench.TOPK, bench.EMBED
    base = syn.synth_db_embeddings(100, n_rows, D).to(dev)
    copies = [base] + [base.clone() for _ in range(NC - 1)]
    eng = OnlineRetrievalEngine(model, base, k=k, max_batch=B, max_tokens=64, depth=depth, scan_ctas=scan_ctas,
                                lstm_clusters=lstm_clusters)
    lib = eng.lib
    batches = [syn.synth_queries(1000 + i, B) for i in range(4)]
    d_text = []
    for b in batches:
        eng.vocab.stage_texts(b, eng.h_stage)
        d_text.append(eng.h_stage.to(dev, copy=True))
    others = torch.nn.functional.normalize(torch.rand(R * B, D, device=dev), dim=1)
    slots = []
    for sl in range(depth):
        s = {}
        s["q_all"] = others.clone()
        s["loc"] = torch.empty(2, R * B, k, dtype=torch.int64, device=dev)
        s["mine"] = torch.zeros(2, R, B, k, dtype=torch.int64, device=dev)
        s["final"] = torch.zeros(2, B, k, dtype=torch.int64, device=dev)
        s["ws"] = torch.empty(max(256, lib.t2p_retrieve_topk_workspace(R * B, n_rows, D, k)), dtype=torch.uint8, device=dev)
        slots.append(s)
        eng.slots[sl].d_stage.copy_(d_text[sl % 4])

    stamps = torch.zeros(depth, 8, dtype=torch.int64, device=dev)
    log = torch.zeros(4096, 8, dtype=torch.int64, device=dev)   # [step][stamp]: copied out of `stamps` at the end of every step
    counter = [0]

    def mark(sl, j):
        if timeline:
            stamp_lib.stamp(ctypes.c_void_p(stamps[sl, j:].data_ptr()), ctypes.c_void_p(_lib.stream_ptr(dev)))

    def step(sl, db):
        s, es = slots[sl], eng.slots[sl]
        st = _lib.stream_ptr(dev)
        mark(sl, 0)
        if only in (None, "lstm"):
            eng.enqueue_tokenize(sl)
            mark(sl, 1)
            eng.enqueue_encode(slot=sl)
            mark(sl, 2)
        if only in (None, "topk"):
            s["q_all"][:B].copy_(es.q)
            _lib.check(lib.t2p_retrieve_topk_ex(s["q_all"].data_ptr(), db.data_ptr(), R * B, n_rows, D, k, 0,
                                                eng.db_norm2_max.data_ptr(), eng.topk_flags, s["loc"][0].data_ptr(),
                                                s["loc"][1].data_ptr(), eng.stats.data_ptr(), s["ws"].data_ptr(), s["ws"].numel(), st),
                       "topk")
            s["mine"][:, 0].copy_(s["loc"][:, :B])
            _lib.check(lib.t2p_topk_merge(s["mine"][0].data_ptr(), s["mine"][1].data_ptr(), R, B, k, k, s["final"][0].data_ptr(),
                                          s["final"][1].data_ptr(), st), "merge")
            mark(sl, 3)

    graphs = [[None] * NC for _ in range(depth)]
    for sl in range(depth):
        with torch.cuda.stream(eng.slots[sl].stream):
            step(sl, copies[0])
    torch.cuda.synchronize()
    for sl in range(depth):
        for c in range(NC):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step(sl, copies[c])
            graphs[sl][c] = g
    torch.cuda.synchronize()
    main = torch.cuda.current_stream()

    C = _lib.C
    execs, streams = (C.c_void_p * K)(), (C.c_void_p * K)()
    for i in range(K):
        execs[i] = graphs[i % depth][i % NC].raw_cuda_graph_exec()
        streams[i] = eng.slots[i % depth].stream.cuda_stream

    def region(ev0, ev1):
        ev0.record()
        if native:
            _lib.check(lib.t2p_serving_replay_many(execs, streams, K, _lib.stream_ptr(dev), 1), "replay_many")
            ev1.record()
            return
        for sl in range(depth):
            eng.slots[sl].stream.wait_event(ev0)
        for i in range(K):
            with torch.cuda.stream(eng.slots[i % depth].stream):
                graphs[i % depth][i % NC].replay()
        for sl in range(depth):
            main.wait_stream(eng.slots[sl].stream)
        ev1.record()

    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        region(a, b)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(200)]
    for a, b in evs:
        region(a, b)
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b) for a, b in evs])
    # serial latency of one step
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(eng.slots[0].stream):
        a.record()
        for i in range(50):
            graphs[0][i % NC].replay()
        b.record()
    torch.cuda.synchronize()
    if timeline:
        # one region, the stamps of every step copied out right after it (each slot runs at most ceil(K/depth) steps)
        ev0 = torch.cuda.Event()
        t_ref = torch.zeros(8, dtype=torch.int64, device=dev)
        stamp_lib.stamp(ctypes.c_void_p(t_ref.data_ptr()), ctypes.c_void_p(_lib.stream_ptr(dev)))
        ev0.record()
        for sl in range(depth):
            eng.slots[sl].stream.wait_event(ev0)
        for i in range(K):
            with torch.cuda.stream(eng.slots[i % depth].stream):
                graphs[i % depth][i % NC].replay()
                log[i].copy_(stamps[i % depth], non_blocking=True)
        torch.cuda.synchronize()
        t0 = int(t_ref[0].item())
        rows = (log[:K].cpu().numpy() - t0) / 1e3
        print("step slot   start    +tok   +lstm   +topk   (us since region start; durations)")
        for i in range(K):
            r = rows[i]
            print(f"{i:4d} {i % depth:4d} {r[0]:8.1f} {r[1]-r[0]:7.1f} {r[2]-r[1]:7.1f} {r[3]-r[2]:7.1f}   end {r[3]:8.1f}")
    host_us = None
    if native:
        import time
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            _lib.check(lib.t2p_serving_replay_many(execs, streams, K, _lib.stream_ptr(dev), 1), "replay_many")
        host_us = (time.perf_counter() - t0) / (20 * K) * 1e6
        torch.cuda.synchronize()
    return {"R": R, "native": native, "host_us_per_launch": host_us, "ncopies": NC, "depth": depth, "K": K, "scan_ctas": scan_ctas, "lstm_clusters": lstm_clusters, "only": only,
            "us_per_step": round(float(np.median(t)) / K * 1e3, 2), "region_ms": round(float(np.median(t)), 4),
            "serial_us": round(a.elapsed_time(b) / 50 * 1e3, 1), "q_per_s_per_gpu": round(B * K / (float(np.median(t)) * 1e-3))}


if __name__ == "__main__":
    spec = sys.argv[1:] or ["8:12:20:40"]
    for sp in spec:
        f = sp.split(":")
        R, depth, K, ctas = int(f[0]), int(f[1]), int(f[2]), int(f[3])
        lc = int(f[4]) if len(f) > 4 and f[4] else None
        only = (f[5] or None) if len(f) > 5 else None
        tl = len(f) > 6 and "t" in f[6]
        nat = len(f) > 6 and "n" in f[6]
        nc = int(f[7]) if len(f) > 7 else None
        print(json.dumps(run(R, depth, K, ctas, lc, n_rows=10000 if R == 1 else 12500, only=only, timeline=tl, native=nat, ncopies=nc)), flush=True)
