"""GPU: all-pairs scores + top-k through the C ABI vs the float64 oracle / the reference's golden vectors."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden
from text2pos_cvpr2022_b200 import synthetic as syn
from text2pos_cvpr2022_b200.retrieval import CellDatabase, retrieve_topk, topk_merge, shard_bounds

pytestmark = pytest.mark.gpu


def _check(db, q, k, idx_base=0):
    idx, sc = retrieve_topk(q.cuda(), db.cuda(), k, idx_base)
    ref_i, ref_s = oracle.retrieval.topk(db.numpy(), q.numpy(), min(k, db.shape[0]))
    kk = ref_i.shape[1]
    np.testing.assert_array_equal(idx.cpu().numpy()[:, :kk], ref_i + idx_base)  # bit-exact indices
    np.testing.assert_allclose(sc.cpu().numpy()[:, :kk], ref_s, rtol=1e-12, atol=1e-15)
    if kk < k:
        assert (idx.cpu().numpy()[:, kk:] == -1).all() and np.isneginf(sc.cpu().numpy()[:, kk:]).all()


def test_golden_reference_loop_vectors():
    z, metas = load_golden("retrieval.npz")
    for i, m in enumerate(metas):
        db = syn.synth_db_embeddings(m["db_seed"], m["N"], m["D"])
        q = syn.synth_query_embeddings(m["q_seed"], m["Q"], m["D"])
        idx, _ = retrieve_topk(q.cuda(), db.cuda(), m["k"])
        np.testing.assert_array_equal(idx.cpu().numpy(), z[f"top{i}"])


@pytest.mark.parametrize("B,N,D,k", [(1, 128, 256, 10), (64, 10000, 256, 10), (64, 12500, 256, 10), (3, 37, 32, 5),
                                      (70, 1000, 256, 10), (5, 300, 100, 1), (2, 4, 256, 10), (9, 2047, 130, 26),
                                      (64, 40000, 256, 10), (17, 777, 384, 10)])
def test_topk_matches_oracle(B, N, D, k):
    _check(syn.synth_db_embeddings(N + D, N, D), syn.synth_query_embeddings(B + 1, B, D), k, idx_base=0)


def test_idx_base_and_ties():
    db = syn.synth_db_embeddings(5, 500, 64)
    db[100] = db[7]
    db[300] = db[7]
    db[301] = db[7]  # exact duplicates: ties must resolve to ascending index
    q = torch.cat([db[7:8], syn.synth_query_embeddings(6, 3, 64)])
    _check(db, q, 10, idx_base=12500)
    idx, _ = retrieve_topk(q.cuda(), db.cuda(), 4)
    assert idx[0].tolist() == [7, 100, 300, 301]


def test_sharded_merge_equals_unsharded_on_device():
    db = syn.synth_db_embeddings(8, 100000, 256)
    q = syn.synth_query_embeddings(9, 64, 256).cuda()
    full_i, full_s = retrieve_topk(q, db.cuda(), 10)
    li, ls = [], []
    for lo, hi in shard_bounds(100000, 8):
        i, s = retrieve_topk(q, db[lo:hi].cuda(), 10, idx_base=lo)
        li.append(i)
        ls.append(s)
    mi, ms = topk_merge(torch.stack(ls), torch.stack(li), 10)
    assert torch.equal(mi, full_i) and torch.equal(ms, full_s)
    ref_i, _ = oracle.retrieval.topk(db.numpy(), q.cpu().numpy(), 10)
    np.testing.assert_array_equal(full_i.cpu().numpy(), ref_i)


def test_full_size_properties():
    """Config-3-sized DB on one GPU: size-independent properties (sortedness, self-retrieval, idempotence)."""
    db = syn.synth_db_embeddings(10, 100000, 256).cuda()
    q = db[torch.arange(0, 100000, 1571)[:64]].clone()
    cdb = CellDatabase(db, cell_ids=[f"0010_{i:05d}" for i in range(100000)])
    idx, sc = cdb.topk(q, 10)
    assert (idx[:, 0].cpu() == torch.arange(0, 100000, 1571)[:64]).all()  # a row retrieves itself first
    assert (sc[:, 1:] <= sc[:, :-1]).all()
    idx2, sc2 = cdb.topk(q, 10)
    assert torch.equal(idx, idx2) and torch.equal(sc, sc2)
    ids = cdb.topk_ids(q[:2], 3)
    assert ids.shape == (2, 3) and ids[0, 0] == "0010_00000"
