"""Timeline of CTA (0,0) of retrieve_scan_tc_kernel from the trace build (tools/make_scan_trace.py).
   usage: python tools/diag_scan_trace.py B N cap"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from text2pos_cvpr2022_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "bin", "libexp_SCT.so")
from text2pos_cvpr2022_b200 import synthetic as syn
from text2pos_cvpr2022_b200.retrieval import retrieve_topk, db_row_norm2_max
B, N, cap = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
db = syn.synth_db_embeddings(1, N, 256).cuda(); q = syn.synth_query_embeddings(2, B, 256).cuda()
nm = db_row_norm2_max(db)
lib = _lib.load()
lib.t2p_debug_scan_trace.argtypes = [ctypes.c_void_p]
for it in range(3):
    retrieve_topk(q, db, 10, 0, norm2_max=nm, flags=_lib.retrieve_max_ctas(cap))
    torch.cuda.synchronize()
buf = np.zeros(128, dtype=np.uint64)
lib.t2p_debug_scan_trace(buf.ctypes.data)
t0 = int(buf[0]); rel = lambda i: (int(buf[i]) - t0) / 1e3 if buf[i] else float("nan")
print(f"setup done {rel(1):.2f}  Q in TMEM {rel(2):.2f}  MMA thread past q_full {rel(3):.2f}  epilogue done {rel(4):.2f}  kernel end {rel(5):.2f} us")
for lt in range(32):
    if buf[8 + lt]:
        print(f"tile {lt:2d}: MMAs issued {rel(8+lt):7.2f}  accumulator ready {rel(40+lt):7.2f}  scanned {rel(72+lt):7.2f}")
if buf[96]:
    print("tile 2, per K chunk: [before wait, after wait, after 4 MMAs + commit] us")
    for kc in range(8):
        print(f"  kc {kc}: {rel(96+3*kc):7.2f} {rel(97+3*kc):7.2f} {rel(98+3*kc):7.2f}")
