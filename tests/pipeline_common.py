"""Shared fixtures of the pipeline tests: the synthetic KITTI360Pose-shaped scene of BASELINE config 5 (small), the model
weights, and the reference-style padding factory / transform (numpy global RNG) used when the restated pipeline is compared
with the unmodified reference functions."""
import types

import numpy as np

from text2pos_cvpr2022_b200 import default_args, synthetic as syn

ARGS = dict(top_k=[1, 3, 5], threshs=[5, 10, 15], pad_size=16, batch_size=8, ranking_loss="pairwise", coarse_oracle=False,
            coarse_random=False, street_oracle=False, use_test_set=False)


def pipeline_args(**kw):
    d = dict(ARGS)
    d.update(kw)
    return types.SimpleNamespace(**d)


def scene(seed=11, n_cells=14, n_poses=6):
    """Two scenes interleaved on a 10 m grid (cells overlap like KITTI360Pose), up to 20 objects per cell so that the top-k
    dataset both cuts and pads."""
    ds = syn.SynthCoarseDataset(seed, n_cells, n_poses, scenes=("0010", "0003"), grid_stride=10.0, max_objects=20)
    return ds, syn.SynthLoader(ds, batch_size=4)


def coarse_state_dict(seed=5):
    from text2pos_cvpr2022_b200.cell_retrieval import CellRetrievalNetwork

    m = CellRetrievalNetwork(syn.KNOWN_CLASSES, syn.COLOR_NAMES, syn.known_words(), default_args(embed_dim=256))
    syn.randomize_module_(m, seed, gain=2.0)
    return m, {k: v.detach().clone() for k, v in m.state_dict().items()}


def fine_state_dict(seed=7):
    """Random fine-matcher weights in a regime where matches actually occur: with plain random weights every object encoding
    collapses onto one direction (cosine 0.997 between objects) and the assignment is uniform, so nothing would ever be
    matched and the pose head would never run.  Large-gain encoders + a negative shift before the last ReLU of the merge MLP
    keep the encodings apart (cosine ~0.4-0.7); a x6 final projection sharpens the scores (several matches per cell)."""
    from text2pos_cvpr2022_b200.superglue_matcher import SuperGlueMatch

    m = SuperGlueMatch(syn.KNOWN_CLASSES, syn.COLOR_NAMES, syn.known_words(), default_args(embed_dim=128, num_layers=6))
    spec = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    sd = syn.synth_state_dict(spec, seed, gain=0.4)
    syn.superglue_peaky_(sd, "superglue.", scale=5.0)
    big = syn.synth_state_dict(spec, seed, gain=8.0)
    for k in sd:
        if k.startswith("object_encoder.") or k.startswith("language_encoder."):
            sd[k] = big[k]
    sd["object_encoder.mlp_merge.0.1.bias"] = sd["object_encoder.mlp_merge.0.1.bias"] - 1.0
    sd["superglue.final_proj.weight"] = sd["superglue.final_proj.weight"] * 6.0
    m.load_state_dict(sd)
    m.eval()
    return m, sd


def reference_padding():
    """``Object3d.create_padding`` (datapreparation/kitti360pose/imports.py:74-83) on numpy's GLOBAL RNG, as a duck type."""
    return syn.SynthObject3d(-1, np.random.rand(8, 3) * 0.001, np.zeros((8, 3)), "pad")


def reference_transform():
    """``T.Compose([T.FixedPoints(256), T.NormalizeScale()])`` (evaluation/pipeline.py:290-293) from whatever torch_geometric
    resolves (the real package, or shims/torch_geometric)."""
    from text2pos_cvpr2022_b200 import compat

    compat.install()
    import torch_geometric.transforms as T

    return T.Compose([T.FixedPoints(256), T.NormalizeScale()])


class RecordingModel:
    """Wraps a fine model and keeps every output (the reference's run_fine returns accuracies only)."""

    def __init__(self, model):
        self.model, self.outputs = model, []

    def __call__(self, objects, hints, object_points):
        out = self.model(objects, hints, object_points)
        self.outputs.append({k: np.asarray(v.detach().cpu().numpy()) for k, v in out.items()})
        return out
