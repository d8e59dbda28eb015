"""Host cost of the end-to-end serving call per batch: time spent inside submit() and collect() (diagnostic)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from text2pos_cvpr2022_b200 import synthetic as syn
from text2pos_cvpr2022_b200.serving import OnlineRetrievalEngine

dev = torch.device("cuda", 0)
model = bench.build_model().to(dev)
db = syn.synth_db_embeddings(100, 10000, bench.EMBED).to(dev)
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 20
eng = OnlineRetrievalEngine(model, db, k=10, max_batch=64, max_tokens=64, depth=depth)
eng.capture_all("g")
batches = [syn.synth_queries(1000 + i, 64) for i in range(4)]
n = 4000
for rep in range(2):
    ts = tc = 0.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        if len(eng._inflight) == depth:
            a = time.perf_counter(); eng.collect(); tc += time.perf_counter() - a
        a = time.perf_counter(); eng.submit(batches[i % 4], graph_key="g"); ts += time.perf_counter() - a
    while eng._inflight:
        eng.collect()
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
    print(f"depth {depth}: {tot / n * 1e6:.2f} us per batch; in submit {ts / n * 1e6:.2f}, in collect {tc / n * 1e6:.2f} (incl. ~0.1 us of timer calls)")
b = batches[0]
t0 = time.perf_counter()
for _ in range(20000):
    blob = ("\0".join(b) + "\0").encode("utf-8")
print(f"join+encode: {(time.perf_counter() - t0) / 20000 * 1e6:.2f} us for {len(blob)} bytes")
s = eng.slots[0]
t0 = time.perf_counter()
for _ in range(20000):
    eng._check_counts(s)
print(f"_check_counts: {(time.perf_counter() - t0) / 20000 * 1e6:.2f} us")
