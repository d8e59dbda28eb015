#!/usr/bin/env python
"""Diagnosis: per-role clock64 trace of sa_edge_tc_kernel (experimental SAT_TRACE build of the library)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from text2pos_cvpr2022_b200 import _lib
_lib.LIB_PATH = os.environ["T2P_DIAG_LIB"]
from text2pos_cvpr2022_b200 import default_args, synthetic as syn
from text2pos_cvpr2022_b200.cell_retrieval import CellRetrievalNetwork

model = CellRetrievalNetwork(syn.KNOWN_CLASSES, syn.COLOR_NAMES, syn.known_words(), default_args(embed_dim=256))
syn.randomize_module_(model, 5, gain=2.0)
model = model.eval().to("cuda")
packed = syn.synth_packed_cells(0, 256).to("cuda")
for _ in range(2):
    model.encode_cells_packed(packed)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_longlong * (16 * 256))()
lib.t2p_debug_trace.argtypes = [C.c_void_p]
print("rc", lib.t2p_debug_trace(buf))
t = np.array(buf, dtype=np.int64).reshape(16, 256)
names = {0: "B0 builder got free table", 1: "B1 table published", 2: "M2 mma sees table", 3: "M3 mma has accumulator", 4: "M4 mma sees full stage",
         5: "M5 mma committed", 6: "P6 producer sees table", 7: "P7 producer has stage", 8: "P8 producer filled", 9: "P9 producer arrived",
         10: "E10 epi sees table", 11: "E11 epi sees accumulator", 12: "E12 epi done"}
for lo, hi, what in ((30, 250, os.environ.get("T2P_TRACE_WHAT", "traced instantiation")),):
    print("==", what, "items", lo, hi)
    seg = t[:, lo:hi]
    for ev in range(13):
        col = seg[ev][seg[ev] != 0]
        per = np.diff(col).astype(np.float64)
        if per.size:
            print(f"  {names[ev]:32s} period between recorded items: median {np.median(per):8.0f} cycles, mean {per.mean():8.0f}")
    chain = [(0, 1), (1, 6), (6, 7), (7, 8), (8, 9), (9, 4), (4, 5), (5, 11), (10, 11), (11, 12), (1, 12)]
    for a, b in chain:
        ok = (seg[a] != 0) & (seg[b] != 0)
        d = (seg[b] - seg[a])[ok].astype(np.float64)
        if d.size:
            print(f"  {names[a][:3]} -> {names[b][:3]}: median {np.median(d):8.0f}  mean {d.mean():8.0f}  max {d.max():8.0f}")
