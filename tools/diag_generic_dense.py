"""Does the fp32 CUDA-core scan (generic path) rank rows right when the scores near the top are denser than fp32 resolves? (diagnostic)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, oracle
from text2pos_cvpr2022_b200 import _lib
from text2pos_cvpr2022_b200.retrieval import retrieve_topk
g = torch.Generator().manual_seed(11)
bad = 0
for D, noise in ((256, 0.004), (256, 0.0004), (96, 0.0004), (260, 0.0002), (512, 0.0002)):
    q = torch.nn.functional.normalize(torch.randn(4, D, generator=g))
    db = torch.nn.functional.normalize(q[0:1] + noise * torch.randn(6000, D, generator=g))
    db[17] = db[4000]
    idx, sc = retrieve_topk(q.cuda(), db.cuda(), 10, 0, flags=_lib.RETRIEVE_FORCE_GENERIC)
    ri, rs = oracle.retrieval.topk(db.numpy(), q.numpy(), 10)
    ok = np.array_equal(idx.cpu().numpy(), ri)
    print(D, noise, "identical" if ok else f"MISMATCH rows {np.where((idx.cpu().numpy() != ri).any(1))[0].tolist()}")
    bad += 0 if ok else 1
print("bad", bad)
