"""GPU: the online engine (strings in, top-k out) -- native tokeniser + staging copies + kernels (+ CUDA graph)."""
import numpy as np
import pytest
import torch

import oracle
from conftest import cpu_state_dict
from text2pos_cvpr2022_b200 import synthetic as syn
from text2pos_cvpr2022_b200.serving import OnlineRetrievalEngine

pytestmark = pytest.mark.gpu


def test_engine_query_matches_oracle_and_graph_replay(coarse_model):
    B, N, k = 16, 3000, 10
    db = syn.synth_db_embeddings(21, N, 256)
    eng = OnlineRetrievalEngine(coarse_model, db.cuda(), k=k, max_batch=B, max_tokens=64, idx_base=100)
    texts = syn.synth_queries(5, B)
    idx, sc = eng.query(texts)
    idx, sc = idx.copy(), sc.copy()
    sd = cpu_state_dict(coarse_model)
    with torch.no_grad():
        q_ref = oracle.text.encode_text(sd, texts, coarse_model.language_encoder.known_words)
    # the engine's own text embeddings agree with the oracle to 1e-4 and rank identically on this well separated DB
    np.testing.assert_allclose(eng.q.cpu().numpy(), q_ref.numpy(), atol=1e-4, rtol=0)
    ref_i, ref_s = oracle.retrieval.topk(db.numpy(), eng.q.cpu().numpy(), k)
    np.testing.assert_array_equal(idx - 100, ref_i)
    np.testing.assert_allclose(sc, ref_s, rtol=1e-12, atol=1e-15)
    eng.capture("g")
    idx2, sc2 = eng.query(texts, graph_key="g")
    np.testing.assert_array_equal(idx2, idx)
    np.testing.assert_array_equal(sc2, sc)
    other = syn.synth_queries(6, B)
    idx3, _ = eng.query(other, graph_key="g")
    idx4, _ = eng.query(other)
    np.testing.assert_array_equal(idx3.copy(), idx4)
    assert eng.stats.cpu().tolist()[1] == 0  # nothing needed the exact rescan


def test_engine_unicode_and_errors(coarse_model):
    db = syn.synth_db_embeddings(22, 500, 256)
    eng = OnlineRetrievalEngine(coarse_model, db.cuda(), k=5, max_batch=2, max_tokens=32)
    a = ["The pose is north of a gray building.", "Die Pose ist südlich einer grünen Wand."]  # 2nd: OOV words, non-ASCII
    idx, _ = eng.query(a)
    want = coarse_model.encode_text(a)
    np.testing.assert_allclose(eng.q.cpu().numpy(), want.cpu().numpy(), atol=1e-6, rtol=0)
    with pytest.raises(ValueError):
        eng.query(["only one"])
    with pytest.raises(ValueError):
        eng.query(["", "x"])


def test_engine_pipelined_submit_collect_equals_query(coarse_model):
    """depth-2 engine: results of submit()/collect() with two batches in flight equal the synchronous query()."""
    B, N, k = 8, 2000, 10
    db = syn.synth_db_embeddings(23, N, 256).cuda()
    ref = OnlineRetrievalEngine(coarse_model, db, k=k, max_batch=B, max_tokens=64)
    eng = OnlineRetrievalEngine(coarse_model, db, k=k, max_batch=B, max_tokens=64, depth=2)
    eng.capture_all("g")
    batches = [syn.synth_queries(30 + i, B) for i in range(5)]
    want = []
    for b in batches:
        i, s = ref.query(b)
        want.append((i.copy(), s.copy()))
    got = []
    for n, b in enumerate(batches):
        if len(eng._inflight) == 2:
            i, s = eng.collect()
            got.append((i.copy(), s.copy()))
        eng.submit(b, graph_key="g" if n % 2 else None)
    while eng._inflight:
        i, s = eng.collect()
        got.append((i.copy(), s.copy()))
    assert len(got) == len(want)
    for (gi, gs), (wi, ws) in zip(got, want):
        np.testing.assert_array_equal(gi, wi)
        np.testing.assert_array_equal(gs, ws)
    with pytest.raises(RuntimeError):
        eng.submit(batches[0]), eng.submit(batches[1]), eng.submit(batches[2])
    while eng._inflight:
        eng.collect()
    i0, _ = eng.query(batches[0])  # the synchronous call still works on a pipelined engine
    np.testing.assert_array_equal(i0, want[0][0])


def test_replay_many_runs_every_captured_step_on_its_slot(coarse_model):
    """Device-resident loop (``t2p_serving_replay_many``): K captured steps launched by one native call, forked from and joined
    into the current stream -- every slot ends up with the result of the text staged in it, for both captured DBs."""
    B, N, k, depth = 8, 2000, 10, 3
    db_a = syn.synth_db_embeddings(24, N, 256).cuda()
    db_b = syn.synth_db_embeddings(25, N, 256).cuda()
    db_b = db_b * (db_a.norm(dim=1).max() / db_b.norm(dim=1).max())  # alternative copies share the norm bound of the engine's DB
    ref = OnlineRetrievalEngine(coarse_model, db_a, k=k, max_batch=B, max_tokens=64)
    eng = OnlineRetrievalEngine(coarse_model, db_a, k=k, max_batch=B, max_tokens=64, depth=depth)
    batches = [syn.synth_queries(40 + i, B) for i in range(depth)]
    for sl, b in enumerate(batches):  # stage the text of slot sl once (resident from now on)
        used, ascii_ = eng.vocab.stage_texts(b, eng.slots[sl].h_stage)
        assert ascii_
        eng.slots[sl].d_stage.copy_(eng.slots[sl].h_stage)
    eng.capture_all("a", db_a)
    eng.capture_all("b", db_b)
    for key, db in (("a", db_a), ("b", db_b)):
        for sl in range(depth):
            eng.slots[sl].out_idx.fill_(-7)
        plan = eng.replay_plan([key] * (2 * depth))  # every slot twice
        ev = torch.cuda.Event()
        eng.replay_many(plan)
        ev.record()          # recorded on the current stream AFTER the join: covers all slots
        ev.synchronize()
        ref.set_db(db)
        for sl, b in enumerate(batches):
            want_i, want_s = ref.query(b)
            np.testing.assert_array_equal(eng.slots[sl].out_idx.cpu().numpy(), want_i)
            np.testing.assert_array_equal(eng.slots[sl].out_scores.cpu().numpy(), want_s)
    with pytest.raises(KeyError):
        eng.replay_plan(["missing"])


def test_device_tokenizer_matches_the_python_rules():
    """t2p_tokenize_device == models/modules.py:60-72 (remove '.' ',', lower, split on whitespace, OOV -> 0, zero pad) on
    templated hints and on edge cases: punctuation inside words, runs of separators, empty and over-long descriptions."""
    from text2pos_cvpr2022_b200 import _lib
    from text2pos_cvpr2022_b200.modules import tokenize

    kw = {w: i + 1 for i, w in enumerate(syn.known_words())}
    kw["<unk>"] = 0
    vocab = _lib.Vocab(kw)
    vocab.to_device("cuda:0")
    T = 70
    texts = syn.synth_queries(3, 40) + [
        "Hello, World.  The POSE\tis north,of a gray building .", "x", "a.b,c d", " lead and trail \n", "north\x1csouth\x0bwest",
        "building" * 30 + " north", "The pose is north of a gray-building.", "north " * 70,
    ]
    empty_rows = [len(texts), len(texts) + 1]
    texts += ["", ".,.,  ,"]
    long_row = len(texts)
    texts += ["north " * 80]  # 80 tokens > T
    n = len(texts)
    cap = _lib.load().t2p_stage_texts_capacity(n, sum(len(t) + 1 for t in texts))
    h_stage = torch.zeros(cap, dtype=torch.uint8).pin_memory()
    d_stage = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    used, ascii_ = vocab.stage_texts(texts, h_stage, d_stage)
    assert ascii_ and used <= cap
    tok = torch.full((n, T), -7, dtype=torch.int32, device="cuda")
    ln = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    vocab.tokenize_device(d_stage, n, tok, ln)
    torch.cuda.synchronize()
    rt, rl = tokenize(texts[:long_row], kw)
    got_l = ln.cpu().numpy()
    np.testing.assert_array_equal(got_l[:long_row], rl)
    assert got_l[long_row] == T + 1 and all(got_l[r] == 0 for r in empty_rows)
    got = tok.cpu().numpy()
    np.testing.assert_array_equal(got[:long_row, : rt.shape[1]], rt)
    assert (got[:long_row, rt.shape[1]:] == 0).all()
    # non-ASCII batches are flagged for the host tokeniser and not copied
    _, ascii2 = vocab.stage_texts(["süd of a wall"], h_stage, d_stage)
    assert not ascii2
