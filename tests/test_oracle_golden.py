"""CPU: the oracle against golden vectors produced by the reference itself (tests/golden/make_golden.py)."""
import numpy as np
import torch

import oracle
from conftest import load_golden
from text2pos_cvpr2022_b200 import synthetic as syn


def test_get_mlp_matches_reference():
    z, metas = load_golden("get_mlp.npz")
    for i, m in enumerate(metas):
        sd = syn.synth_state_dict([(k, s) for k, s in m["spec"]], m["seed"])
        x = torch.randn(17, m["channels"][0], generator=torch.Generator().manual_seed(m["xseed"]))
        y = oracle.mlp.get_mlp(sd, "", x)
        np.testing.assert_allclose(y.numpy(), z[f"y{i}"], rtol=1e-5, atol=1e-6)
        assert oracle.mlp.mlp_channels(sd, "") == m["channels"]


def _superglue_case(name):
    z, m = load_golden(f"superglue_{name}.npz")
    sd = syn.synth_state_dict([(k, s) for k, s in m["spec"]], m["seed"], m["gain"])
    if m["peaky"]:
        sd = syn.superglue_peaky_(sd, scale=m["peaky"])
    d0, d1 = syn.synth_descriptor_pairs(m["seed"] + 1, m["B"], m["M"], m["N"], m["D"])
    out = oracle.superglue.superglue_forward(sd, "", d0, d1, m["num_layers"], m["iters"])
    return z, m, out


def test_superglue_matches_reference():
    for name in ("fine", "small"):
        z, m, out = _superglue_case(name)
        np.testing.assert_allclose(out["P"].numpy(), z["P"], rtol=2e-4, atol=2e-6)
        np.testing.assert_array_equal(out["matches0"].numpy(), z["matches0"])
        np.testing.assert_array_equal(out["matches1"].numpy(), z["matches1"])
        np.testing.assert_allclose(out["matching_scores0"].numpy(), z["matching_scores0"], rtol=2e-4, atol=2e-6)
        np.testing.assert_allclose(out["matching_scores1"].numpy(), z["matching_scores1"], rtol=2e-4, atol=2e-6)


def test_superglue_known_answers():
    """Derived KATs: Sinkhorn marginals, mutual matches."""
    z, m, out = _superglue_case("fine")
    P = out["P"]
    M, N = m["M"], m["N"]
    np.testing.assert_allclose(P[:, :M, :].sum(2).numpy(), 1.0, atol=2e-3)
    np.testing.assert_allclose(P[:, :, :N].sum(1).numpy(), 1.0, atol=2e-3)
    np.testing.assert_allclose(P[:, M, :].sum(1).numpy(), N, atol=2e-2)
    np.testing.assert_allclose(P[:, :, N].sum(1).numpy(), M, atol=2e-2)
    m0, m1 = out["matches0"], out["matches1"]
    assert (m0 >= 0).sum() >= 16  # the planted correspondences are found
    for b in range(m0.shape[0]):
        for i, j in enumerate(m0[b].tolist()):
            if j >= 0:
                assert m1[b, j] == i and P[b, i, j] > 0.2


def test_language_encoder_matches_reference():
    for name in ("coarse", "fine"):
        z, m = load_golden(f"language_encoder_{name}.npz")
        sd = syn.synth_state_dict([(k, s) for k, s in m["spec"]], m["seed"])
        assert m["words"] == syn.known_words()
        kw = {w: i + 1 for i, w in enumerate(m["words"])}
        kw["<unk>"] = 0
        tokens, lengths = oracle.text.tokenize(m["texts"], kw)
        assert lengths.min() != lengths.max()  # ragged
        enc = oracle.text.language_encoder(sd, "", tokens, lengths)
        np.testing.assert_allclose(enc.numpy(), z["encodings"], rtol=1e-4, atol=2e-6)


def test_retrieval_matches_reference_loop():
    z, metas = load_golden("retrieval.npz")
    for i, m in enumerate(metas):
        db = syn.synth_db_embeddings(m["db_seed"], m["N"], m["D"]).numpy()
        q = syn.synth_query_embeddings(m["q_seed"], m["Q"], m["D"]).numpy()
        idx, sc = oracle.retrieval.topk(db, q, m["k"])
        np.testing.assert_array_equal(idx, z[f"top{i}"])
        assert (np.diff(sc, axis=1) <= 0).all()
        if m["N"] <= 1000:
            np.testing.assert_array_equal(oracle.retrieval.reference_loop(db, q, m["k"]), z[f"top{i}"])


def test_retrieval_shard_merge_equals_unsharded():
    db = syn.synth_db_embeddings(3, 999, 64).numpy()
    q = syn.synth_query_embeddings(4, 9, 64).numpy()
    idx, sc = oracle.retrieval.topk(db, q, 10)
    bounds = [0, 100, 450, 451, 999]
    si, ss = [], []
    for a, b in zip(bounds[:-1], bounds[1:]):
        i_, s_ = oracle.retrieval.topk(db[a:b], q, min(10, b - a))
        si.append(i_ + a)
        ss.append(s_)
    mi, ms = oracle.retrieval.merge_shards(si, ss, 10)
    np.testing.assert_array_equal(mi, idx)
    np.testing.assert_allclose(ms, sc, rtol=1e-13, atol=0)


def test_retrieval_tie_rule():
    db = np.zeros((6, 4), dtype=np.float32)
    db[:, 0] = [0.5, 1.0, 1.0, 0.25, 1.0, 0.5]
    q = np.array([[1, 0, 0, 0]], dtype=np.float32)
    idx, _ = oracle.retrieval.topk(db, q, 5)
    assert idx.tolist() == [[1, 2, 4, 0, 5]]  # score desc, index asc


def test_reference_port_language_encoder_matches_reference():
    from oracle.reference_port import LanguageEncoderPort

    for name in ("coarse", "fine"):
        z, m = load_golden(f"language_encoder_{name}.npz")
        sd = syn.synth_state_dict([(k, s) for k, s in m["spec"]], m["seed"])
        kw = {w: i + 1 for i, w in enumerate(m["words"])}
        kw["<unk>"] = 0
        enc = LanguageEncoderPort(sd, "", kw)
        np.testing.assert_allclose(enc(m["texts"]).numpy(), z["encodings"], rtol=1e-5, atol=1e-6)
