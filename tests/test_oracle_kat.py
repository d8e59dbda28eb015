"""CPU: derived known-answer tests for the parity-unpinned pieces of the oracle (SURVEY 8c v, vi)."""
import numpy as np
import torch

import oracle
from oracle.pointnet import ball_query, fps, radius_sq, sa_edges, sqdist_f32
from text2pos_cvpr2022_b200 import synthetic as syn


def test_fps_greedy_property_and_ties():
    cells = syn.synth_packed_cells(1, 3)
    pos = cells.pos.numpy()
    idx = fps(pos, 128)
    assert (idx[:, 0] == 0).all()
    for o in range(0, pos.shape[0], 7):
        sel = [0]
        for s in range(1, 16):
            d = np.min(sqdist_f32(pos[o][:, None, :], pos[o][None, sel, :]), axis=1)
            assert d[idx[o, s]] == d.max() and idx[o, s] == int(np.argmax(d))  # farthest, lowest index on ties
            sel.append(int(idx[o, s]))
    dup = np.zeros((1, 8, 3), dtype=np.float32)  # all points identical: every tie resolves to index 0
    assert fps(dup, 4).tolist() == [[0, 0, 0, 0]]


def test_ball_query_properties():
    cells = syn.synth_packed_cells(2, 2)
    pos = cells.pos.numpy()
    idx = fps(pos, 128)
    cpos = np.take_along_axis(pos, idx[:, :, None], axis=1)
    nbr, cnt = ball_query(pos, cpos, 0.2)
    r2 = radius_sq(0.2)
    assert cnt.max() == 32 and cnt.min() >= 1
    for o in range(pos.shape[0]):
        for c in range(0, 128, 17):
            n = nbr[o, c, : cnt[o, c]]
            assert (np.diff(n) > 0).all()  # ascending candidate order
            assert (sqdist_f32(pos[o, n], cpos[o, c][None]) < r2).all()
            assert (nbr[o, c, cnt[o, c] :] == -1).all()
            if cnt[o, c] < 32:  # not truncated: the list is the complete ball
                assert cnt[o, c] == int((sqdist_f32(pos[o], cpos[o, c][None]) < r2).sum())


def test_self_loop_quirk_edges_micro_case():
    """2 objects x 4 points, 2 centres each: flat point i feeds flat centre i (hand-derived)."""
    nbr = -np.ones((2, 2, 32), dtype=np.int32)
    nbr[0, 0, :2] = [0, 1]
    nbr[0, 1, :1] = [3]
    nbr[1, 0, :1] = [2]
    nbr[1, 1, :2] = [0, 3]
    cnt = (nbr >= 0).sum(-1).astype(np.int32)
    so, sp, do, dc = sa_edges(nbr, cnt, P=4, self_loop_quirk=True)
    edges = sorted(zip(so.tolist(), sp.tolist(), do.tolist(), dc.tolist()))
    expect = sorted([
        (0, 1, 0, 0), (0, 3, 0, 1), (1, 2, 1, 0), (1, 0, 1, 1), (1, 3, 1, 1),  # radius edges, (0,0)->(0,0) removed
        (0, 0, 0, 0), (0, 1, 0, 1), (0, 2, 1, 0), (0, 3, 1, 1),  # flat i -> flat i, i < 4 centres
    ])
    assert edges == expect
    so, sp, do, dc = sa_edges(nbr, cnt, P=4, self_loop_quirk=False)
    assert len(so) == 6


def test_sa_layer_identity_micro_case():
    """One SA layer with identity-like weights: out = max over edges of relu([x_j, pos_j - pos_i])."""
    sd = {}
    eye = torch.eye(6)
    for i in range(2):
        sd[f"sa.point_conv.local_nn.{i}.0.weight"] = eye.clone()
        sd[f"sa.point_conv.local_nn.{i}.0.bias"] = torch.zeros(6)
        sd[f"sa.point_conv.local_nn.{i}.1.weight"] = torch.ones(6) * float(np.sqrt(1 + 1e-5))
        sd[f"sa.point_conv.local_nn.{i}.1.bias"] = torch.zeros(6)
        sd[f"sa.point_conv.local_nn.{i}.1.running_mean"] = torch.zeros(6)
        sd[f"sa.point_conv.local_nn.{i}.1.running_var"] = torch.ones(6)
    pos = torch.tensor([[[0.0, 0, 0], [0.1, 0, 0], [0.9, 0, 0], [0.95, 0, 0]]])
    x = torch.tensor([[[1.0, 0, 0], [0, 2.0, 0], [0, 0, 3.0], [4.0, 0, 0]]])
    out, cpos, idx, nbr, cnt = oracle.pointnet.set_abstraction(sd, "sa.", x, pos, 0.2, self_loop_quirk=False)
    assert idx.tolist() == [[0, 3]]  # start 0, farthest = point 3
    assert nbr[0, 0, :2].tolist() == [0, 1] and nbr[0, 1, :2].tolist() == [2, 3]
    np.testing.assert_allclose(out[0, 0].numpy(), [1, 2, 0, 0.1, 0, 0], atol=1e-6)
    np.testing.assert_allclose(out[0, 1].numpy(), [4, 0, 3, 0, 0, 0], atol=1e-6)


def test_cell_encoder_invariants():
    from text2pos_cvpr2022_b200 import default_args
    from text2pos_cvpr2022_b200.cell_retrieval import CellRetrievalNetwork

    m = CellRetrievalNetwork(syn.KNOWN_CLASSES, syn.COLOR_NAMES, syn.known_words(), default_args(embed_dim=64))
    syn.randomize_module_(m, 2, gain=2.0)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    cells = syn.synth_packed_cells(3, 4)
    sl = cells.cell_slices()
    with torch.no_grad():
        out = oracle.cells.encode_objects(sd, [cells.rgb[a:b] for a, b in sl], [cells.pos[a:b] for a, b in sl], cells.centers, cells.mean_rgb)
        # a cell's embedding does not depend on which other cells share the call
        one = oracle.cells.encode_objects(sd, [cells.rgb[sl[2][0]:sl[2][1]]], [cells.pos[sl[2][0]:sl[2][1]]],
                                          cells.centers[sl[2][0]:sl[2][1]], cells.mean_rgb[sl[2][0]:sl[2][1]])
    np.testing.assert_allclose(out.norm(dim=1).numpy(), 1.0, atol=1e-5)
    assert (out >= 0).all()  # post-ReLU lin
    np.testing.assert_allclose(out[2].numpy(), one[0].numpy(), atol=1e-6)
    e = torch.nn.functional.normalize(torch.randn(11, 64, generator=torch.Generator().manual_seed(0)))
    knn = oracle.cells.knn_in_cell(e.numpy())
    assert knn.shape == (11, 8) and (knn[:, 0] == np.arange(11)).all()  # self is the nearest
    assert oracle.cells.knn_in_cell(e[:5].numpy()).shape == (5, 5)
