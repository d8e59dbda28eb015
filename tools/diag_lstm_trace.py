"""Per-group timeline of four steps of lstm_tc_kernel (CTA 0) from the trace build (tools/make_lstm_trace.py), throughput
geometry (one cluster per direction, four groups of 16 sequences)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from text2pos_cvpr2022_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "bin", "libexp_LTR.so")
import bench
from text2pos_cvpr2022_b200 import synthetic as syn
from text2pos_cvpr2022_b200.serving import OnlineRetrievalEngine
dev = torch.device("cuda", 0)
model = bench.build_model().to(dev)
db = syn.synth_db_embeddings(100, 10000, bench.EMBED).to(dev)
eng = OnlineRetrievalEngine(model, db, k=10, max_batch=64, max_tokens=64, depth=8)  # depth >= 8: one cluster per direction
lib = eng.lib
lib.t2p_debug_lstm_trace.argtypes = [ctypes.c_void_p]
b = syn.synth_queries(1000, 64)
for _ in range(3):
    eng.query(b)
torch.cuda.synchronize()
buf = np.zeros(4 * 4 * 8, dtype=np.uint64)
lib.t2p_debug_lstm_trace(buf.ctypes.data)
t = buf.reshape(4, 4, 8).astype(np.int64)
t0 = t[0, :, 0].min()
names = ["h arrived (control)", "MMAs issued", "acc ready (epi)", "tmem loaded", "staged", "set barrier", "copy issued"]
print("us since the first event; rows = (step, group)")
print("step g  " + "  ".join(f"{n:>18s}" for n in names))
for s in range(4):
    for g in range(4):
        print(f"{20 + s:4d} {g}  " + "  ".join(f"{(t[s, g, e] - t0) / 1e3:18.2f}" if t[s, g, e] else f"{'-':>18s}" for e in range(7)))
