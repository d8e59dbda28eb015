"""GPU: text encoder (tensor-core / register / shared-memory LSTM kernels) and SuperGlue head vs the reference's golden vectors and the oracle."""
import numpy as np
import pytest
import torch

import oracle
from conftest import cpu_state_dict, load_golden
from text2pos_cvpr2022_b200 import default_args, synthetic as syn
from text2pos_cvpr2022_b200.modules import LanguageEncoder
from text2pos_cvpr2022_b200.superglue import SuperGlue

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["coarse", "fine"])
def test_language_encoder_golden(name):
    z, m = load_golden(f"language_encoder_{name}.npz")
    enc = LanguageEncoder(m["words"], m["D"], bi_dir=True)
    enc.load_state_dict(syn.synth_state_dict([(k, s) for k, s in m["spec"]], m["seed"]))
    enc = enc.cuda().eval()
    out = enc(m["texts"])
    np.testing.assert_allclose(out.cpu().numpy(), z["encodings"], atol=1e-5, rtol=1e-4)


@pytest.mark.parametrize("D,B", [(256, 64), (128, 60), (64, 3), (32, 9), (16, 1)])
def test_language_encoder_vs_oracle(D, B):
    enc = LanguageEncoder(syn.known_words(), D, bi_dir=True)
    syn.randomize_module_(enc, D)
    sd = cpu_state_dict(enc)
    enc = enc.cuda().eval()
    texts = syn.synth_queries(D + 1, B, 6 if D >= 128 else 2)
    texts[0] = "north"  # length 1
    out = enc(texts)
    tokens, lengths = oracle.text.tokenize(texts, enc.known_words)
    ref = oracle.text.language_encoder(sd, "", tokens, lengths)
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), atol=1e-5, rtol=1e-4)


@pytest.mark.parametrize("path,B", [(3, 64), (3, 5), (3, 17), (2, 64), (2, 5), (1, 64)])
def test_language_encoder_every_kernel_path_H256(path, B):
    """H = 256: 1 = shared-memory cluster kernel, 2 = register-resident kernel, 3 = tcgen05 tensor-core kernel (default)."""
    enc = LanguageEncoder(syn.known_words(), 256, bi_dir=True)
    syn.randomize_module_(enc, 77)
    sd = cpu_state_dict(enc)
    enc = enc.cuda().eval()
    texts = syn.synth_queries(300 + B, B, 6)
    texts[0] = "north"  # length 1
    if B > 2:
        texts[2] = "the pose is north of a gray building " * 7  # longest row by far: 56 tokens
    _, desc = enc.t2p_packed()
    desc.path = path
    try:
        out = enc(texts)
    finally:
        desc.path = 0
    tokens, lengths = oracle.text.tokenize(texts, enc.known_words)
    ref = oracle.text.language_encoder(sd, "", tokens, lengths)
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), atol=1e-5, rtol=1e-4)


@pytest.mark.parametrize("max_groups,B", [(2, 64), (1, 64), (1, 32), (1, 19), (3, 64), (2, 130), (4, 64), (1, 17), (2, 3)])
def test_language_encoder_tensor_core_cluster_geometries(max_groups, B):
    """tcgen05 kernel, every launch geometry: clusters of <= 16 sequences (one group), clusters of 17..32 sequences run as
    two ping-pong groups (uneven halves, different longest rows per group), and waves of clusters for large batches."""
    enc = LanguageEncoder(syn.known_words(), 256, bi_dir=True)
    syn.randomize_module_(enc, 78)
    sd = cpu_state_dict(enc)
    enc = enc.cuda().eval()
    texts = syn.synth_queries(400 + B, B, 6)
    texts[0] = "north"  # length 1
    if B > 20:
        texts[20] = "the pose is north of a gray building " * 7  # 56 tokens: lands in the second group of a 32-wide cluster
    texts[B - 1] = "west of a green wall"
    _, desc = enc.t2p_packed()
    desc.path, desc.max_groups = 3, max_groups
    try:
        out = enc(texts)
    finally:
        desc.path, desc.max_groups = 0, 0
    tokens, lengths = oracle.text.tokenize(texts, enc.known_words)
    ref = oracle.text.language_encoder(sd, "", tokens, lengths)
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), atol=1e-5, rtol=1e-4)


def test_encode_text_normalised(coarse_model):
    texts = syn.synth_queries(5, 64)
    out = coarse_model.encode_text(texts)
    ref = oracle.text.encode_text(cpu_state_dict(coarse_model), texts, coarse_model.language_encoder.known_words)
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), atol=1e-5, rtol=1e-4)
    assert torch.allclose(out.norm(dim=1), torch.ones(64, device="cuda"), atol=1e-5)


@pytest.mark.parametrize("name", ["fine", "small"])
def test_superglue_golden(name):
    z, m = load_golden(f"superglue_{name}.npz")
    sg = SuperGlue({"descriptor_dim": m["D"], "GNN_layers": ["self", "cross"] * m["num_layers"],
                    "sinkhorn_iterations": m["iters"], "match_threshold": 0.2})
    sd = syn.synth_state_dict([(k, s) for k, s in m["spec"]], m["seed"], m["gain"])
    if m["peaky"]:
        sd = syn.superglue_peaky_(sd, scale=m["peaky"])
    sg.load_state_dict(sd)
    sg = sg.cuda().eval()
    d0, d1 = syn.synth_descriptor_pairs(m["seed"] + 1, m["B"], m["M"], m["N"], m["D"])
    out = sg(d0.transpose(1, 2).contiguous().cuda(), d1.transpose(1, 2).contiguous().cuda())  # channel-first API
    np.testing.assert_allclose(out["P"].cpu().numpy(), z["P"], rtol=2e-3, atol=1e-5)
    np.testing.assert_array_equal(out["matches0"].cpu().numpy(), z["matches0"])
    np.testing.assert_array_equal(out["matches1"].cpu().numpy(), z["matches1"])
    np.testing.assert_allclose(out["matching_scores0"].cpu().numpy(), z["matching_scores0"], rtol=2e-3, atol=1e-5)
    np.testing.assert_allclose(out["matching_scores1"].cpu().numpy(), z["matching_scores1"], rtol=2e-3, atol=1e-5)
    assert out["matches0"].dtype == torch.int64


@pytest.mark.parametrize("B,M,N,D,L", [(32, 16, 6, 128, 6), (3, 7, 3, 32, 2), (2, 1, 1, 16, 1), (4, 20, 9, 64, 1), (2, 16, 6, 256, 1)])
def test_superglue_vs_oracle(B, M, N, D, L):
    sg = SuperGlue({"descriptor_dim": D, "GNN_layers": ["self", "cross"] * L, "sinkhorn_iterations": 50})
    sd = syn.synth_state_dict([(k, tuple(v.shape)) for k, v in sg.state_dict().items()], B + M, 0.4)
    syn.superglue_peaky_(sd, scale=5.0)
    sg.load_state_dict(sd)
    sg = sg.cuda().eval()
    d0, d1 = syn.synth_descriptor_pairs(B, B, M, N, D)
    out = sg.match_rows(d0.cuda(), d1.cuda(), return_scores=True)
    ref = oracle.superglue.superglue_forward(sd, "", d0, d1, L, 50)
    np.testing.assert_allclose(out["scores"].cpu().numpy(), ref["scores"].numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(out["P"].cpu().numpy(), ref["P"].numpy(), rtol=2e-3, atol=1e-5)
    np.testing.assert_array_equal(out["matches0"].cpu().numpy(), ref["matches0"].numpy())
    np.testing.assert_array_equal(out["matches1"].cpu().numpy(), ref["matches1"].numpy())
    # Sinkhorn marginals hold at any size
    P = out["P"]
    assert torch.allclose(P[:, :M, :].sum(2), torch.ones(B, M, device="cuda"), atol=5e-3)
    assert torch.allclose(P[:, :, :N].sum(1), torch.ones(B, N, device="cuda"), atol=5e-3)


def test_fine_matcher_forward(fine_model):
    """Config 4 through the drop-in API: B=32 samples x 16 (padded) objects x 6 hints, D=128."""
    sd = cpu_state_dict(fine_model)
    rng = np.random.default_rng(4)
    B = 32
    objects, points = [], []
    for b in range(B):
        n_real = int(rng.integers(6, 17))
        objs = [syn.synth_object(rng, obj_id=i) for i in range(n_real)]
        objs += [syn.SynthObject3d.create_padding(rng) for _ in range(16 - n_real)]
        objects.append(objs)
        points.append(syn.batch_object_points(objs, rng))
    hints = syn.synth_hints(9, B)
    out = fine_model(objects, hints, points)
    assert set(out.keys()) == {"P", "matches0", "matches1", "offsets", "matching_scores0", "matching_scores1"}
    assert out.P.shape == (B, 17, 7) and out.matches0.shape == (B, 16) and out.offsets.shape == (B, 6, 2)
    packed = syn.pack_cells(objects, points)
    sl = packed.cell_slices()
    with torch.no_grad():
        obj_ref = oracle.cells.object_encoder(sd, "object_encoder.", [packed.rgb[a:b] for a, b in sl],
                                              [packed.pos[a:b] for a, b in sl], packed.centers, packed.mean_rgb)
        kw = fine_model.language_encoder.known_words
        hint_ref = torch.stack([oracle.text.language_encoder(sd, "language_encoder.", *oracle.text.tokenize(h, kw)) for h in hints])
        ref = oracle.superglue.superglue_match_forward(sd, hint_ref, obj_ref.reshape(B, 16, 128), 6, 50)
    np.testing.assert_allclose(out.offsets.cpu().numpy(), ref["offsets"].numpy(), atol=1e-4, rtol=1e-4)
    np.testing.assert_allclose(out.P.cpu().numpy(), ref["P"].numpy(), rtol=5e-3, atol=1e-4)
    np.testing.assert_array_equal(out.matches0.cpu().numpy(), ref["matches0"].numpy())
    np.testing.assert_array_equal(out.matches1.cpu().numpy(), ref["matches1"].numpy())
