"""CPU: host-side logic of the drop-in modules (tokeniser, packing/BN folding, cell packing, state_dict layout, loud failure)."""
import numpy as np
import pytest
import torch

import oracle
from conftest import cpu_state_dict, load_golden
from text2pos_cvpr2022_b200 import default_args, packing, synthetic as syn
from text2pos_cvpr2022_b200.modules import LanguageEncoder, get_mlp, tokenize
from text2pos_cvpr2022_b200.object_encoder import cell_chunks, obj_cell_start_from_offsets
from text2pos_cvpr2022_b200.retrieval import shard_bounds
from text2pos_cvpr2022_b200.runtime import AttrDict
from text2pos_cvpr2022_b200.superglue import SuperGlue


def _blob_linear(blob, d, x):
    W = blob[d.w_off : d.w_off + d.k * d.n].reshape(d.k, d.n)
    b = blob[d.b_off : d.b_off + d.n]
    return x @ W + b


def test_tokenizer_matches_oracle():
    kw = {w: i + 1 for i, w in enumerate(syn.known_words())}
    kw["<unk>"] = 0
    texts = syn.synth_queries(3, 9) + ["The pose is, north. of a UNKNOWN thing"]
    t0, l0 = oracle.text.tokenize(texts, kw)
    t1, l1 = tokenize(texts, kw)
    np.testing.assert_array_equal(t0, t1)
    np.testing.assert_array_equal(l0, l1)
    assert t1.dtype == np.int32 and 48 <= l1[0] <= 60


def test_bn_folding_matches_oracle_get_mlp():
    m = get_mlp([7, 12, 5])
    syn.randomize_module_(m, 3)
    sd = cpu_state_dict(m)
    bb = packing.BlobBuilder()
    d0 = packing.pack_mlp_layer(bb, sd, "0.")
    d1 = packing.pack_mlp_layer(bb, sd, "1.")
    blob = bb.finish().double().numpy()
    x = torch.randn(11, 7, generator=torch.Generator().manual_seed(0))
    h = np.maximum(_blob_linear(blob, d0, x.double().numpy()), 0)
    y = np.maximum(_blob_linear(blob, d1, h), 0)
    np.testing.assert_allclose(y, oracle.mlp.get_mlp(sd, "", x).numpy(), rtol=1e-5, atol=1e-6)
    assert d0.w_off % 4 == 0 and d1.w_off % 4 == 0  # float4-aligned weight tiles


def test_lstm_input_projection_folding():
    enc = LanguageEncoder(syn.known_words(), 32, bi_dir=True)
    syn.randomize_module_(enc, 4)
    sd = cpu_state_dict(enc)
    bb = packing.BlobBuilder()
    d = packing.pack_lstm(bb, sd, "")
    blob = bb.finish()
    V, H = d.vocab, d.hidden
    xproj = blob[d.xproj_off : d.xproj_off + 2 * V * 4 * H].reshape(2, V, 4 * H)
    emb = sd["word_embedding.weight"]
    ref = emb @ sd["lstm.weight_ih_l0_reverse"].t() + sd["lstm.bias_ih_l0_reverse"] + sd["lstm.bias_hh_l0_reverse"]
    np.testing.assert_allclose(xproj[1].numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)
    whh = blob[d.whh_off : d.whh_off + 2 * H * 4 * H].reshape(2, H, 4 * H)
    np.testing.assert_array_equal(whh[0].numpy(), sd["lstm.weight_hh_l0"].t().numpy())


def test_edgeconv_split_is_equivalent():
    from text2pos_cvpr2022_b200.cell_retrieval import CellRetrievalNetwork

    m = CellRetrievalNetwork(syn.KNOWN_CLASSES, syn.COLOR_NAMES, syn.known_words(), default_args(embed_dim=16))
    syn.randomize_module_(m, 9)
    sd = cpu_state_dict(m)
    bb = packing.BlobBuilder()
    d = packing.pack_cell_aggregation(bb, sd, 16)
    blob = bb.finish().double().numpy()
    g = torch.Generator().manual_seed(1)
    xi, xj = torch.randn(5, 16, generator=g), torch.randn(5, 16, generator=g)
    ab_i = _blob_linear(blob, d.edge_ab, xi.double().numpy())
    ab_j = _blob_linear(blob, d.edge_ab, xj.double().numpy())
    h1 = np.maximum(ab_i[:, :16] + ab_j[:, 16:], 0)
    y = np.maximum(_blob_linear(blob, d.edge_l2, h1), 0)
    ref = oracle.mlp.get_mlp(sd, "graph1.nn.", torch.cat([xi, xj - xi], dim=1))
    np.testing.assert_allclose(y, ref.numpy(), rtol=1e-5, atol=1e-5)


def test_state_dict_layout_matches_reference_modules():
    """Key names AND shapes equal those of the reference classes (recorded in the golden fixtures)."""
    _, m = load_golden("superglue_fine.npz")
    sg = SuperGlue({"descriptor_dim": m["D"], "GNN_layers": ["self", "cross"] * m["num_layers"], "sinkhorn_iterations": m["iters"]})
    assert [(k, list(v.shape)) for k, v in sg.state_dict().items()] == [(k, list(s)) for k, s in m["spec"]]
    _, m = load_golden("language_encoder_coarse.npz")
    le = LanguageEncoder(m["words"], m["D"], bi_dir=True)
    assert [(k, list(v.shape)) for k, v in le.state_dict().items()] == [(k, list(s)) for k, s in m["spec"]]


def test_pack_cells_and_chunking():
    cells, objects, points = syn.synth_cells(0, 5)
    packed = syn.pack_cells(objects, points)
    n = sum(len(o) for o in objects)
    assert packed.pos.shape == (n, 256, 3) and packed.cell_offsets.tolist()[-1] == n
    assert float(packed.pos.abs().max()) < 1.0
    np.testing.assert_allclose(packed.centers[0].numpy(), objects[0][0].get_center(), rtol=1e-6)
    start = obj_cell_start_from_offsets(packed.cell_offsets)
    assert start.tolist() == [o for a, b in packed.cell_slices() for o in [a] * (b - a)]
    off = [0, 5, 9, 20, 21]
    assert cell_chunks(off, 10) == [(0, 2), (2, 3), (3, 4)]
    assert cell_chunks(off, 1000) == [(0, 4)]
    bad = syn.PointBatch(points[0].x[:-1], points[0].pos[:-1], points[0].batch[:-1])
    with pytest.raises(ValueError):
        syn.pack_cells(objects[:1], [bad])


def test_shard_bounds():
    assert shard_bounds(100000, 8) == [(i * 12500, (i + 1) * 12500) for i in range(8)]
    assert shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]


def test_attrdict_and_loud_cpu_failure(coarse_model):
    a = AttrDict()
    a.P = 1
    assert a["P"] == 1 and list(a) == ["P"] and "P" in a
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            coarse_model.encode_text(["The pose is north of a gray box."])
        with pytest.raises(Exception):
            coarse_model.forward()


def test_lstm_register_tiling_is_a_permutation_of_whh():
    """packing._lstm_register_tiling: thread (w, kp, jj) of cluster rank r holds W_hh[g*H + 32r + 4w + jj][4*(8i+kp)+e]."""
    from text2pos_cvpr2022_b200 import packing

    rng = np.random.default_rng(0)
    for H in (32, 64, 128, 256):
        whh_t = rng.standard_normal((2, H, 4 * H))  # W_hh^T per direction
        tiled = packing._lstm_register_tiling(whh_t)
        assert tiled.shape == (2, H // 32, H // 32, 4, 256, 4)
        assert np.array_equal(np.sort(tiled.reshape(2, -1), axis=1), np.sort(whh_t.reshape(2, -1), axis=1))
        for (d, r, i, e, w, kp, jj, g) in [(0, 0, 0, 0, 0, 0, 0, 0), (1, H // 32 - 1, H // 32 - 1, 3, 7, 7, 3, 3), (1, 0, H // 64, 2, 5, 3, 1, 2)]:
            t = w * 32 + kp * 4 + jj
            assert tiled[d, r, i, e, t, g] == whh_t[d, 4 * (8 * i + kp) + e, g * H + 32 * r + 4 * w + jj]


def test_native_tokenizer_matches_the_python_rules():
    """t2p_tokenize (host code of the C ABI) == the reference's rules (models/modules.py:60-72) on ASCII input."""
    import torch

    from text2pos_cvpr2022_b200 import _lib, synthetic as syn
    from text2pos_cvpr2022_b200.modules import tokenize

    kw = {w: i + 1 for i, w in enumerate(syn.known_words())}
    kw["<unk>"] = 0
    vocab = _lib.Vocab(kw)
    texts = syn.synth_queries(3, 40) + [
        "Hello, World.  The POSE\tis north,of a gray building .", "x", ".,.,", "a.b,c d", " lead and trail \n", "north\x1csouth\x0bwest",
    ]
    tok = torch.full((len(texts), 70), -7, dtype=torch.int32)
    ln = torch.zeros(len(texts), dtype=torch.int32)
    longest = vocab.tokenize_into(texts, tok, ln)
    rt, rl = tokenize(texts, kw)
    assert longest == int(rl.max())
    np.testing.assert_array_equal(ln.numpy(), rl)
    np.testing.assert_array_equal(tok[:, : rt.shape[1]].numpy(), rt)
    assert (tok[:, rt.shape[1]:] == 0).all()  # rows are fully written (zero padded)
    small = torch.zeros(1, 3, dtype=torch.int32)
    with pytest.raises(RuntimeError):
        vocab.tokenize_into(["one two three four"], small, torch.zeros(1, dtype=torch.int32))


def test_lstm_tensor_core_images_layout_and_split():
    """packing._lstm_tc_images: fp16 hi/lo split of 2^8*W_hh in the tensor-memory load order [k-unit][row][8 fp16]."""
    from text2pos_cvpr2022_b200 import packing

    rng = np.random.default_rng(1)
    H = 256
    whh_t = rng.uniform(-0.0625, 0.0625, size=(2, H, 4 * H))
    img = packing._lstm_tc_images(whh_t).view(np.float16).reshape(2, 8, 2, 32, 128, 8)
    for (d, r, m, k) in [(0, 0, 0, 0), (1, 7, 127, 255), (0, 3, 77, 130), (1, 5, 9, 63), (0, 2, 64, 64)]:
        unit, gate = m // 4, m % 4
        w = whh_t[d, k, gate * H + 32 * r + unit] * 256.0
        ku, e = k // 8, k % 8
        hi = float(img[d, r, 0, ku, m, e])
        lo = float(img[d, r, 1, ku, m, e])
        assert hi == float(np.float16(w))
        assert abs(hi + lo - w) <= 2.0 ** -21 * abs(w) + 1e-12
    # every weight appears exactly once per part
    hi_sorted = np.sort(img[:, :, 0].astype(np.float64).reshape(2, -1), axis=1)
    want = np.sort((whh_t * 256.0).astype(np.float16).astype(np.float64).reshape(2, -1), axis=1)
    assert np.array_equal(hi_sorted, want)


def test_stage_texts_layout():
    """t2p_stage_texts (host half of the device tokeniser): [int32 offsets[n+1] | pad to 16 | NUL-terminated strings]."""
    import torch

    from text2pos_cvpr2022_b200 import _lib

    vocab = _lib.Vocab({"a": 1})
    texts = ["The pose is north of a gray building.", "", "x y"]
    lib = _lib.load()
    cap = lib.t2p_stage_texts_capacity(len(texts), sum(len(t) + 1 for t in texts))
    assert cap % 16 == 0
    buf = torch.zeros(cap, dtype=torch.uint8)
    used, ascii_ = vocab.stage_texts(texts, buf)
    assert ascii_ and used == cap
    off = buf[:16].view(torch.int32).tolist()
    assert off == [0, 38, 39, 43]
    raw = bytes(buf[16: 16 + 43].tolist())
    assert raw == b"The pose is north of a gray building.\0\0x y\0"
    _, ascii2 = vocab.stage_texts(["grün"], buf)
    assert not ascii2
    with pytest.raises(RuntimeError):
        vocab.stage_texts(["x" * 100], torch.zeros(32, dtype=torch.uint8))


def test_sa_tensor_core_images_layout_and_split():
    """packing._sa_tc_images: fp16 hi/lo split of 2^8*W2 in the 128-byte-swizzled K-major UMMA B layout, per 64-wide K chunk."""
    rng = np.random.default_rng(2)
    for C in (128, 256):
        w = rng.uniform(-0.2, 0.2, size=(C, C))  # [K, N] as stored in the blob
        img = packing._sa_tc_images(w).view(np.float16).reshape(C // 64, 2, C, 64)
        for (k, n) in [(0, 0), (C - 1, C - 1), (70, 5), (63, 9), (64, 127)]:
            chunk, lu, e = k // 64, (k % 64) // 8, k % 8
            pu = lu ^ (n & 7)
            v = w[k, n] * 256.0
            hi, lo = float(img[chunk, 0, n, pu * 8 + e]), float(img[chunk, 1, n, pu * 8 + e])
            assert hi == float(np.float16(v))
            assert abs(hi + lo - v) <= 2.0 ** -21 * abs(v) + 1e-12
        assert np.array_equal(np.sort(img[:, 0].astype(np.float64).reshape(-1)),
                              np.sort((w * 256.0).astype(np.float16).astype(np.float64).reshape(-1)))


def test_pointnet_pack_emits_tensor_core_images():
    """pack_pointnet2: every second local_nn layer (SA1 32 -> 64 zero-padded to [64, 128], SA2 128 -> 128, SA3 256 -> 256 as two
    128-column blocks), the second global-abstraction layer (512 -> 1024, eight 128-wide column blocks) and the x parts of the
    dense layers get fp16 hi/lo images.  Image (k, n) of column block j equals the folded weight the fp32 kernels read from the
    same blob."""
    from text2pos_cvpr2022_b200.pointnet2 import PointNet2

    pn = PointNet2(len(syn.KNOWN_CLASSES), len(syn.COLOR_NAMES), default_args(embed_dim=256))
    syn.randomize_module_(pn, 3)
    sd = cpu_state_dict(pn)
    bb = packing.BlobBuilder()
    d = packing.pack_pointnet2(bb, sd, "", True)
    blob = bb.finish().numpy()
    assert d.sa_l2_tc_off[0] >= 0 and d.sa_l2_tc_off[1] >= 0 and d.sa_l2_tc_off[2] >= 0 and d.ga_l2_tc_off >= 0
    assert all(d.dense_tc_off[i] >= 0 for i in range(5)) and d.dense_tc_off[5] == -1
    # SA1: [1 chunk][hi|lo][128 rows (64 real channels)][64 fp16], k >= 32 zero
    w1 = blob[d.sa_l2[0].w_off: d.sa_l2[0].w_off + 32 * 64].astype(np.float64).reshape(32, 64)
    img1 = blob[d.sa_l2_tc_off[0]: d.sa_l2_tc_off[0] + 64 * 128].view(np.uint32).view(np.float16).reshape(1, 2, 128, 64)
    assert float(np.abs(img1[0, :, 64:]).max()) == 0.0  # padded output channels
    for (k, n) in [(0, 0), (31, 63), (7, 20)]:
        pu = (k // 8) ^ (n & 7)
        assert float(img1[0, 0, n, pu * 8 + k % 8]) == float(np.float16(w1[k, n] * 256.0))
    assert float(np.abs(img1[0, :, 5, ((40 // 8) ^ 5) * 8 + 0]).max()) == 0.0  # padded K rows
    # SA3: column block j = n // 128
    w3 = blob[d.sa_l2[2].w_off: d.sa_l2[2].w_off + 256 * 256].astype(np.float64).reshape(256, 256)
    img3 = blob[d.sa_l2_tc_off[2]: d.sa_l2_tc_off[2] + 256 * 256].view(np.uint32).view(np.float16).reshape(2, 4, 2, 128, 64)
    for (k, n) in [(0, 0), (255, 255), (70, 130), (129, 77)]:
        j, nn, chunk, lu, e = n // 128, n % 128, k // 64, (k % 64) // 8, k % 8
        assert float(img3[j, chunk, 0, nn, (lu ^ (nn & 7)) * 8 + e]) == float(np.float16(w3[k, n] * 256.0))
    K, N = d.ga_l2.k, d.ga_l2.n
    assert (K, N) == (512, 1024)
    w = blob[d.ga_l2.w_off: d.ga_l2.w_off + K * N].astype(np.float64).reshape(K, N)
    img = blob[d.ga_l2_tc_off: d.ga_l2_tc_off + K * N].view(np.uint32).view(np.float16).reshape(N // 128, K // 64, 2, 128, 64)
    for (k, n) in [(0, 0), (511, 1023), (70, 300), (129, 777)]:
        j, nn = n // 128, n % 128
        chunk, lu, e = k // 64, (k % 64) // 8, k % 8
        pu = lu ^ (nn & 7)
        v = w[k, n] * 256.0
        hi, lo = float(img[j, chunk, 0, nn, pu * 8 + e]), float(img[j, chunk, 1, nn, pu * 8 + e])
        assert hi == float(np.float16(v)) and abs(hi + lo - v) <= 2.0 ** -21 * abs(v) + 1e-12


def test_weights_outside_the_fp16_range_keep_the_fp32_kernels():
    """BatchNorm folding with a tiny running_var can push a folded weight past the fp16 range (2^8 * |w| >= 65504): such a layer
    must not get tensor-core images (the kernels then take the exact-fp32 path) instead of silently packing inf."""
    from text2pos_cvpr2022_b200.pointnet2 import PointNet2

    pn = PointNet2(len(syn.KNOWN_CLASSES), len(syn.COLOR_NAMES), default_args(embed_dim=256))
    syn.randomize_module_(pn, 3)
    sd = cpu_state_dict(pn)
    sd["sa2.point_conv.local_nn.1.1.running_var"][5] = 0.0  # 1/sqrt(eps) = 316 ...
    sd["sa2.point_conv.local_nn.1.1.weight"][5] = 40.0       # ... times gamma: output channel 5 of SA2's second layer explodes
    d = packing.pack_pointnet2(packing.BlobBuilder(), sd, "", True)
    assert d.sa_l2_tc_off[1] == -1 and d.sa_l2_tc_off[2] >= 0 and d.ga_l2_tc_off >= 0
    assert packing.fits_fp16_split(np.array([200.0]), 256.0) and not packing.fits_fp16_split(np.array([240.0]), 256.0)
    assert not packing.fits_fp16_split(np.array([np.inf]), 1.0)
    enc = LanguageEncoder(syn.known_words(), 256, bi_dir=True)
    syn.randomize_module_(enc, 1)
    sd = cpu_state_dict(enc)
    assert packing.pack_lstm(packing.BlobBuilder(), sd, "").whh_tc_off >= 0
    sd["lstm.weight_hh_l0"][3, 7] = 300.0
    d = packing.pack_lstm(packing.BlobBuilder(), sd, "")
    assert d.whh_tc_off == -1 and d.xproj4_off == -1 and d.whh_reg_off >= 0


def test_superglue_tensor_core_stream_layout():
    """pack_superglue (D = 128): the weight stream of superglue_tc_kernel holds 20 stages per layer + 2 in consumption order;
    q'/k'/v' columns and merge' rows are in head-major order (new channel h*32 + d = reference channel 4 d + h)."""
    from text2pos_cvpr2022_b200.superglue import SuperGlue

    sg = SuperGlue({"descriptor_dim": 128, "GNN_layers": ["self", "cross"], "sinkhorn_iterations": 5})
    syn.randomize_module_(sg, 4, gain=0.5)
    sd = cpu_state_dict(sg)
    bb = packing.BlobBuilder()
    d = packing.pack_superglue(bb, sd, "", ["self", "cross"], 5, 0.2)
    blob = bb.finish().numpy()
    assert d.tc_w_off >= 0 and d.tc_b_off >= 0
    n_stages = 2 * 20 + 2
    st = blob[d.tc_w_off: d.tc_w_off + n_stages * 8192].view(np.uint32).view(np.float16).reshape(n_stages, 2, 128, 64)
    perm = np.array([4 * dd + h for h in range(4) for dd in range(32)])

    def folded(lin):
        return blob[lin.w_off: lin.w_off + lin.k * lin.n].astype(np.float64).reshape(lin.k, lin.n)

    def elem(stage, n, kk):  # value of (output channel n, k = kk within the 64-wide chunk) of a stage
        u, e = kk // 8, kk % 8
        pu = u ^ (n & 7)
        return float(st[stage, 0, n, pu * 8 + e]) + float(st[stage, 1, n, pu * 8 + e])

    L = 1
    base = L * 20
    wq, wm, w0, w3 = folded(d.q[L]), folded(d.merge[L]), folded(d.mlp0[L]), folded(d.mlp3[L])
    for n, k in [(0, 0), (5, 7), (127, 127), (33, 64), (100, 90)]:
        assert abs(elem(base + 0 + k // 64, n, k % 64) - 256.0 * wq[k, perm[n]]) < 1e-3 * abs(256.0 * wq[k, perm[n]]) + 1e-6
        assert abs(elem(base + 6 + k // 64, n, k % 64) - 256.0 * wm[perm[k], n]) < 1e-3 * abs(256.0 * wm[perm[k], n]) + 1e-6
    # mlp0: stages 8..15 = (nb0: kc0, kc1), (nb1: kc0, kc1), (nb0: kc2, kc3), (nb1: kc2, kc3)
    order = [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (0, 3), (1, 2), (1, 3)]
    for i, (nb, kc) in enumerate(order):
        n, kk = 17, 9
        want = 256.0 * w0[kc * 64 + kk, nb * 128 + n]
        assert abs(elem(base + 8 + i, n, kk) - want) < 1e-3 * abs(want) + 1e-6
    for kc in range(4):
        want = 256.0 * w3[kc * 64 + 3, 77]
        assert abs(elem(base + 16 + kc, 77, 3) - want) < 1e-3 * abs(want) + 1e-6
    bias = blob[d.tc_b_off: d.tc_b_off + 2 * 896 + 128]
    bq = blob[d.q[L].b_off: d.q[L].b_off + 128]
    np.testing.assert_array_equal(bias[L * 896: L * 896 + 128], bq[perm])
    np.testing.assert_array_equal(bias[2 * 896:], blob[d.final_proj.b_off: d.final_proj.b_off + 128])
    # D != 128: no tensor-core stream
    sg64 = SuperGlue({"descriptor_dim": 64, "GNN_layers": ["self"], "sinkhorn_iterations": 5})
    d64 = packing.pack_superglue(packing.BlobBuilder(), cpu_state_dict(sg64), "", ["self"], 5, 0.2)
    assert d64.tc_w_off == -1 and d64.tc_b_off == -1
