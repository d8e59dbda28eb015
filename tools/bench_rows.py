#!/usr/bin/env python
"""Per-row measurements of the SURVEY section-8 hot-path rows that bench.py's headline metric does not cover, each with the CPU
oracle (the checker, timed on a bounded sample on the host cores) beside it and a parity check on the sample:

    a2-a5  DB build: cells/s through PointNet++ -> object encoder -> DGCNN -> pool -> lin (packed fast path)
    a8-a10 SuperGlue head, BASELINE config 4 kernel-only variant: B=32, M=16, N=6, D=128, 12 layers, 50 Sinkhorn iterations
    a11    fine matcher through the drop-in API (objects + hints in, matches + offsets out), B=32

    python tools/bench_rows.py [--cells 256] [--iters 20]        # prints one JSON line per row
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def cuda_time(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters  # ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    import oracle
    from text2pos_cvpr2022_b200 import default_args, synthetic as syn
    from text2pos_cvpr2022_b200.cell_retrieval import CellRetrievalNetwork
    from text2pos_cvpr2022_b200.superglue import SuperGlue
    from text2pos_cvpr2022_b200.superglue_matcher import SuperGlueMatch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dev = torch.device("cuda", 0)

    # ---- a2-a5: DB build ------------------------------------------------------------------------------------------------
    model = CellRetrievalNetwork(syn.KNOWN_CLASSES, syn.COLOR_NAMES, syn.known_words(), default_args(embed_dim=256))
    syn.randomize_module_(model, 5, gain=2.0)
    model.eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(dev)
    packed = syn.synth_packed_cells(0, args.cells)
    n_obj = int(packed.pos.shape[0])
    d_packed = packed.to(dev)
    enc = lambda: model.encode_cells_packed(d_packed)
    out = enc()
    ms = cuda_time(enc, args.iters)
    n_cpu = min(8, args.cells)
    sl = packed.cell_slices()[:n_cpu]
    o1 = sl[-1][1]
    t0 = time.perf_counter()
    with torch.no_grad():
        ref = oracle.cells.encode_objects(sd, [packed.rgb[a:b] for a, b in sl], [packed.pos[a:b] for a, b in sl],
                                          packed.centers[:o1], packed.mean_rgb[:o1])
    cpu_s = time.perf_counter() - t0
    err = float((out[:n_cpu].cpu() - ref).abs().max())
    print(json.dumps({"row": "a2-a5 DB build (PointNet++ -> object encoder -> DGCNN -> pool -> lin)", "cells": args.cells,
                      "objects": n_obj, "points_per_object": 256, "ms": ms, "value": args.cells / (ms * 1e-3), "unit": "cells/s",
                      "objects_per_s": n_obj / (ms * 1e-3),
                      "cpu_baseline": {"value": n_cpu / cpu_s, "unit": "cells/s", "cores": cores, "kind": "port",
                                       "sample": f"{n_cpu} cells through the CPU oracle, one pass"},
                      "max_abs_err_vs_oracle": err, "tolerance": 1e-4}), flush=True)

    # ---- a8-a10: SuperGlue head (config 4, kernel-only) -----------------------------------------------------------------
    B, M, N, D, L = 32, 16, 6, 128, 6
    sg = SuperGlue({"descriptor_dim": D, "GNN_layers": ["self", "cross"] * L, "sinkhorn_iterations": 50})
    sgsd = syn.synth_state_dict([(k, tuple(v.shape)) for k, v in sg.state_dict().items()], B + M, 0.4)
    syn.superglue_peaky_(sgsd, scale=5.0)
    sg.load_state_dict(sgsd)
    sg = sg.to(dev).eval()
    d0, d1 = syn.synth_descriptor_pairs(B, B, M, N, D)
    g0, g1 = d0.to(dev), d1.to(dev)
    run = lambda: sg.match_rows(g0, g1)
    o = run()
    ms = cuda_time(run, args.iters * 5)
    t0 = time.perf_counter()
    with torch.no_grad():
        ref = oracle.superglue.superglue_forward(sgsd, "", d0, d1, L, 50)
    cpu_s = time.perf_counter() - t0
    same = bool(np.array_equal(o["matches0"].cpu().numpy(), ref["matches0"].numpy()))
    flops = B * (2 * 12 * (M + N) * (4 * D * D + 2 * D * 2 * D + 2 * D * D) + 4 * 6 * D * (M + N) ** 2)
    print(json.dumps({"row": "a8-a10 SuperGlue head (12 GNN layers, final proj, 50 Sinkhorn iterations, matching)",
                      "B": B, "M": M, "N": N, "D": D, "ms": ms, "value": B / (ms * 1e-3), "unit": "samples/s",
                      "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
                      "cpu_baseline": {"value": B / cpu_s, "unit": "samples/s", "cores": cores, "kind": "port",
                                       "sample": "one batch of 32 through the CPU oracle (pinned to the reference's SuperGlue by the golden vectors)"},
                      "matches_identical_to_oracle": same}), flush=True)

    # ---- a11: fine matcher through the drop-in API ------------------------------------------------------------------------
    fm = SuperGlueMatch(syn.KNOWN_CLASSES, syn.COLOR_NAMES, syn.known_words(), default_args(embed_dim=128, num_layers=6))
    fsd = syn.synth_state_dict([(k, tuple(v.shape)) for k, v in fm.state_dict().items()], 7, gain=0.4)
    syn.superglue_peaky_(fsd, "superglue.", scale=5.0)
    fm.load_state_dict(fsd)
    fm = fm.eval().to(dev)
    rng = np.random.default_rng(4)
    objects, points = [], []
    for b in range(B):
        n_real = int(rng.integers(6, 17))
        objs = [syn.synth_object(rng, obj_id=i) for i in range(n_real)]
        objs += [syn.SynthObject3d.create_padding(rng) for _ in range(16 - n_real)]
        objects.append(objs)
        points.append(syn.batch_object_points(objs, rng))
    hints = syn.synth_hints(9, B)
    run = lambda: fm(objects, hints, points)
    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 5 * 1e3
    print(json.dumps({"row": "a11 SuperGlueMatch.forward through the drop-in API (host lists in: packing + H2D inside the time)",
                      "B": B, "objects_per_cell": 16, "hints": 6, "D": 128, "ms": ms, "value": B / (ms * 1e-3), "unit": "samples/s"}),
          flush=True)


if __name__ == "__main__":
    main()
