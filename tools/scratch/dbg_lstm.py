import os, sys, copy
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from text2pos_cvpr2022_b200 import synthetic as syn
from text2pos_cvpr2022_b200.modules import tokenize
from text2pos_cvpr2022_b200.serving import OnlineRetrievalEngine
dev = torch.device("cuda", 0)
model = bench.build_model().to(dev)
kw = model.language_encoder.known_words
t, l = tokenize(syn.synth_queries(1000, 64), kw)
d_tok = torch.from_numpy(t.astype(np.int32)).to(dev); d_len = torch.from_numpy(l.astype(np.int32)).to(dev)
base = syn.synth_db_embeddings(100, 1000, 256).to(dev)
eng = OnlineRetrievalEngine(model, base, k=10, max_batch=64, max_tokens=t.shape[1])
for groups in (4, 5, 6, 7):
  for dbg in (0,):
    os.environ["T2P_LSTM_DBG"] = str(dbg); os.environ["T2P_LSTM_GROUPS"] = str(groups)
    for _ in range(5): eng.enqueue_encode(d_tok, d_len)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): eng.enqueue_encode(d_tok, d_len)
    e1.record(); torch.cuda.synchronize()
    print(f"groups={groups} dbg={dbg} (1=wait for all sources first): {e0.elapsed_time(e1)*10:.1f} us", flush=True)
