"""GPU: the eval_epoch drop-in (training/coarse.py:68-167) on a synthetic KITTI360Pose-shaped dataset vs the same
evaluation run through the CPU oracle encoders + the reference's float64 numpy retrieval loop."""
import types

import numpy as np
import pytest
import torch

import oracle
from conftest import cpu_state_dict
from text2pos_cvpr2022_b200 import synthetic as syn
from text2pos_cvpr2022_b200.coarse_eval import eval_epoch

pytestmark = pytest.mark.gpu


def test_eval_epoch_matches_oracle_pipeline(coarse_model):
    ds = syn.SynthCoarseDataset(seed=3, n_cells=40, n_poses=70)
    loader = syn.SynthLoader(ds, batch_size=16)
    args = types.SimpleNamespace(top_k=[1, 3, 5], batch_size=16, ranking_loss="pairwise")
    acc, acc_close, retrievals, cell_enc, text_enc = eval_epoch(coarse_model, loader, args, return_encodings=True)
    assert set(acc) == {1, 3, 5} and len(retrievals) == 70 and cell_enc.shape == (40, 256) and text_enc.shape == (70, 256)

    # the same evaluation on the CPU: oracle encoders, then the reference's loop (float64 scores, argsort)
    sd = cpu_state_dict(coarse_model)
    cds = ds.get_cell_dataset()
    items = [cds[i] for i in range(len(cds))]
    packed = syn.pack_cells([it["objects"] for it in items], [it["object_points"] for it in items])
    sl = packed.cell_slices()
    with torch.no_grad():
        ref_cells = oracle.cells.encode_objects(sd, [packed.rgb[a:b] for a, b in sl], [packed.pos[a:b] for a, b in sl],
                                                packed.centers, packed.mean_rgb).numpy()
        ref_text = oracle.text.encode_text(sd, [ds[i]["texts"] for i in range(len(ds))],
                                           coarse_model.language_encoder.known_words).numpy()
    np.testing.assert_allclose(cell_enc, ref_cells, atol=1e-4, rtol=0)
    np.testing.assert_allclose(text_enc, ref_text, atol=1e-4, rtol=0)
    ids = np.array([c.id for c in cds.cells])
    # rank the GPU's own encodings with the reference loop: identical retrievals (the embeddings differ by ~1e-7 from the
    # oracle's, which may swap near-ties, so the loop is fed the same numbers the kernel saw)
    ref_idx = oracle.retrieval.reference_loop(cell_enc.astype(np.float32).astype(np.float64),
                                              text_enc.astype(np.float32).astype(np.float64), 5)
    for q in range(70):
        assert list(retrievals[q]) == list(ids[ref_idx[q]])
    hits = {k: np.mean([ds.all_poses[q].cell_id in ids[ref_idx[q]][:k] for q in range(70)]) for k in (1, 3, 5)}
    assert all(abs(acc[k] - hits[k]) < 1e-12 for k in (1, 3, 5))
    assert all(0.0 <= acc_close[k] <= 1.0 for k in (1, 3, 5)) and acc_close[5] >= acc_close[1]
