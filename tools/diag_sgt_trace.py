"""Phase timeline of two GNN layers of superglue_tc_kernel (CTA 0) from the trace build (tools/make_sgt_trace.py), B=32, M=16, N=6."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from text2pos_cvpr2022_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "bin", "libexp_GTR.so")
from text2pos_cvpr2022_b200 import synthetic as syn
from text2pos_cvpr2022_b200.superglue import SuperGlue
cfg = {"descriptor_dim": 128, "GNN_layers": ["self", "cross"] * 6, "sinkhorn_iterations": 50, "match_threshold": 0.2}
sg = SuperGlue(cfg)
syn.randomize_module_(sg, 3, gain=1.0)
sg = sg.eval().cuda()
g = torch.Generator().manual_seed(1)
d0 = torch.nn.functional.normalize(torch.randn(32, 16, 128, generator=g), dim=-1).cuda()
d1 = torch.nn.functional.normalize(torch.randn(32, 6, 128, generator=g), dim=-1).cuda()
for _ in range(3):
    out = sg.match_rows(d0, d1)
torch.cuda.synchronize()
lib = _lib.load()
lib.t2p_debug_sgt_trace.argtypes = [ctypes.c_void_p]
buf = np.zeros(64, dtype=np.uint64)
lib.t2p_debug_sgt_trace(buf.ctypes.data)
t = buf.astype(np.int64)
names = ["layer start", "A(X) built", "qkv MMAs done", "attention + A(msg)", "merge MMA done", "A(X) built", "MLP0 [X] done", "A(merged) built",
         "MLP0 [merged] done", "A(hidden lo) built", "W3 lo done", "A(hidden hi) built", "W3 hi done"]
for L in range(2):
    base = 16 * L
    print(f"layer {4 + L} ({'self' if L == 0 else 'cross'}): us since layer start, (delta)")
    prev = t[base]
    for i, n in enumerate(names):
        if t[base + i]:
            print(f"   {n:22s} {(t[base + i] - t[base]) / 1e3:7.2f}  (+{(t[base + i] - prev) / 1e3:5.2f})")
            prev = t[base + i]
    nxt = t[base + 16] if L == 0 else t[40]
    print(f"   {'residual, layer end':22s} {(nxt - t[base]) / 1e3:7.2f}  (+{(nxt - prev) / 1e3:5.2f})")
print(f"final projection + scores + Sinkhorn + matching: {(t[41] - t[40]) / 1e3:.2f} us")
