#!/usr/bin/env python
"""Diagnosis: DB build (encode_cells_packed) host time vs device time per call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from text2pos_cvpr2022_b200 import _lib
if os.environ.get("T2P_DIAG_LIB"):  # an experimental build of the library (knock-out experiments)
    _lib.LIB_PATH = os.environ["T2P_DIAG_LIB"]
from text2pos_cvpr2022_b200 import default_args, synthetic as syn
from text2pos_cvpr2022_b200.cell_retrieval import CellRetrievalNetwork

model = CellRetrievalNetwork(syn.KNOWN_CLASSES, syn.COLOR_NAMES, syn.known_words(), default_args(embed_dim=256))
syn.randomize_module_(model, 5, gain=2.0)
model = model.eval().to("cuda")
packed = syn.synth_packed_cells(0, int(sys.argv[1]) if len(sys.argv) > 1 else 256).to("cuda")
for _ in range(3):
    model.encode_cells_packed(packed)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(20):
    model.encode_cells_packed(packed)
e1.record()
host = (time.perf_counter() - t0) / 20 * 1e3
torch.cuda.synchronize()
print(f"host enqueue {host:.3f} ms/call, device {e0.elapsed_time(e1) / 20:.3f} ms/call, objects {packed.pos.shape[0]}")
