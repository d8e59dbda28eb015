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


# ---- tensor-core (tcgen05 + TMA) scan: certified TF32 candidates, float64 re-rank, exact rescan ---------------------
from text2pos_cvpr2022_b200 import _lib  # noqa: E402
from text2pos_cvpr2022_b200.retrieval import db_row_norm2_max  # noqa: E402


def _check_flags(db, q, k, flags, expect_rescans=None, idx_base=0):
    stats = torch.zeros(2, dtype=torch.int32, device="cuda")
    idx, sc = retrieve_topk(q.cuda(), db.cuda(), k, idx_base, flags=flags, stats=stats)
    ref_i, ref_s = oracle.retrieval.topk(db.numpy(), q.numpy(), min(k, db.shape[0]))
    kk = ref_i.shape[1]
    np.testing.assert_array_equal(idx.cpu().numpy()[:, :kk], ref_i + idx_base)
    np.testing.assert_allclose(sc.cpu().numpy()[:, :kk], ref_s, rtol=1e-12, atol=1e-15)
    st = stats.cpu().tolist()
    assert st[0] + st[1] == q.shape[0]  # every query either certified (tensor path or float32 scan) or rescanned exactly
    if expect_rescans is not None:
        assert st[1] == expect_rescans, st
    return st


@pytest.mark.parametrize("B,N,D,k", [(64, 10000, 256, 10), (64, 12500, 256, 10), (1, 128, 256, 10), (130, 3000, 128, 5),
                                      (7, 50, 32, 10), (3, 5, 64, 10), (64, 100000, 256, 10), (20, 2500, 96, 26), (64, 20000, 256, 1),
                                      # several query tiles: DB tile resident, queries streamed (sharded-engine shapes) / DB streamed
                                      (512, 12500, 256, 10), (256, 12500, 256, 10), (300, 777, 64, 10), (200, 40000, 256, 10)])
def test_tensor_core_path_equals_float64_oracle(B, N, D, k):
    db = syn.synth_db_embeddings(N + D + 1, N, D)
    q = syn.synth_query_embeddings(B + 2, B, D)
    st = _check_flags(db, q, k, 0)
    if N >= 1000:  # well separated synthetic scores: the TF32 candidates certify, no rescan needed
        assert st[1] == 0, st
    _check_flags(db, q, k, _lib.RETRIEVE_FORCE_GENERIC)
    # serving geometry: few scan CTAs, each streaming many DB tiles against its resident query tile
    for cap in (40, 8):
        st = _check_flags(db, q, k, _lib.retrieve_max_ctas(cap), idx_base=1000 * cap)
        if N >= 1000:
            assert st[1] == 0, st


def test_tensor_core_forced_rescan_is_exact():
    db = syn.synth_db_embeddings(3, 4000, 256)
    q = syn.synth_query_embeddings(4, 9, 256)
    _check_flags(db, q, 10, _lib.RETRIEVE_FORCE_RESCAN, expect_rescans=9, idx_base=777)


def test_tensor_core_dense_scores_fall_back_to_the_exact_rescan():
    """Rows that differ from the query direction by tiny perturbations: the scores near the top are closer together than
    the TF32 error bound, so TF32 cannot rank them; the certification must notice and the result must still be exact."""
    g = torch.Generator().manual_seed(11)
    q = torch.nn.functional.normalize(torch.randn(4, 256, generator=g))
    db = torch.nn.functional.normalize(q[0:1] + 0.004 * torch.randn(6000, 256, generator=g))
    db[17] = db[4000]  # exact duplicates on top of it
    st = _check_flags(db, q, 10, 0)
    assert st[1] >= 1, st  # at least the query aligned with the cluster needs the rescan


@pytest.mark.parametrize("D,noise", [(256, 0.0004), (96, 0.0004), (260, 0.0002), (512, 0.0002)])
def test_float32_scan_dense_scores_fall_back_to_the_exact_rescan(D, noise):
    """The CUDA-core scan (shapes outside the tensor path, or forced) pre-selects with float32 scores.  Rows that differ from the
    query direction by perturbations below float32 resolution cannot be ranked that way: the certification must notice and the
    float64 rescan must give the reference's ranking (index order among exact duplicates included)."""
    g = torch.Generator().manual_seed(11)
    q = torch.nn.functional.normalize(torch.randn(4, D, generator=g))
    db = torch.nn.functional.normalize(q[0:1] + noise * torch.randn(6000, D, generator=g))
    db[17] = db[4000]
    st = _check_flags(db, q, 10, _lib.RETRIEVE_FORCE_GENERIC)
    assert st[1] >= 1, st  # the query aligned with the cluster
    assert st[0] >= 1, st  # the others certify from the float32 scan


def test_tensor_core_unnormalised_rows_and_negative_scores():
    g = torch.Generator().manual_seed(12)
    db = torch.randn(5000, 64, generator=g) * torch.rand(5000, 1, generator=g) * 3.0  # norms vary, scores of both signs
    q = torch.randn(33, 64, generator=g)
    _check_flags(db, q, 10, 0)
    db_neg = -torch.abs(db)  # every score against a positive query is negative
    _check_flags(db_neg, torch.abs(q), 7, 0)


def test_db_row_norm2_max_kernel():
    db = syn.synth_db_embeddings(5, 3000, 256) * 1.7
    got = float(db_row_norm2_max(db.cuda()).cpu())
    want = float((db.double() ** 2).sum(1).max())
    assert abs(got - want) <= 1e-5 * want


def test_scan_with_eight_epilogue_warps_in_a_fresh_process():
    """T2P_SCAN_EPI=8 (two epilogue warps per TMEM lane quadrant, two key lists per query and CTA) is read once per process:
    run the parity check for a single-tile, a multi-tile and a sharded-engine shape in a subprocess with it set."""
    import os
    import subprocess
    import sys

    code = (
        "import numpy as np, torch, oracle\n"
        "from text2pos_cvpr2022_b200 import _lib, synthetic as syn\n"
        "from text2pos_cvpr2022_b200.retrieval import retrieve_topk\n"
        "for B, N, cap in ((64, 10000, 0), (64, 10000, 40), (512, 12500, 24), (7, 50, 0)):\n"
        "    db = syn.synth_db_embeddings(N + 1, N, 256); q = syn.synth_query_embeddings(B + 2, B, 256)\n"
        "    st = torch.zeros(2, dtype=torch.int32, device='cuda')\n"
        "    idx, sc = retrieve_topk(q.cuda(), db.cuda(), 10, 5, flags=_lib.retrieve_max_ctas(cap), stats=st)\n"
        "    ri, rs = oracle.retrieval.topk(db.numpy(), q.numpy(), 10)\n"
        "    assert np.array_equal(idx.cpu().numpy(), ri + 5), (B, N, cap)\n"
        "    assert np.allclose(sc.cpu().numpy(), rs, rtol=1e-12, atol=1e-15)\n"
        "    assert sum(st.cpu().tolist()) == B\n"
        "print('ok')\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, T2P_SCAN_EPI="8", PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    out = subprocess.run([sys.executable, "-c", code], env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr
