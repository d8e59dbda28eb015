"""Generate golden vectors by RUNNING THE REFERENCE ITSELF (dev container only).

    python tests/golden/make_golden.py            # needs /root/reference, writes tests/golden/*.npz

The reference modules that import cleanly here (``models/superglue.py``,
``models/modules.py``) and the verbatim numpy retrieval loop
(``training/coarse.py:134-140``) are executed on seeded inputs; only the seeds,
key/shape lists and the reference OUTPUTS are stored -- weights and inputs are
re-derived from the seed by ``text2pos_cvpr2022_b200.synthetic`` in the tests, which keeps
the fixtures tiny.  Nothing at test/bench time reads ``/root/reference``.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from text2pos_cvpr2022_b200 import synthetic as syn  # noqa: E402


def _import_reference():
    if not hasattr(np, "int"):
        np.int = int  # models/modules.py:69 uses the removed numpy alias
    # the repo ships its own `models` package (the drop-in shims); load the reference's under another name
    import importlib.util

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    ref_modules = load("ref_models_modules", os.path.join(REF, "models/modules.py"))
    ref_superglue = load("ref_models_superglue", os.path.join(REF, "models/superglue.py"))
    return ref_modules, ref_superglue


def spec_of(module):
    return [(k, list(v.shape)) for k, v in module.state_dict().items()]


def golden_superglue(ref_superglue, name, D, num_layers, iters, B, M, N, seed, gain, peaky):
    cfg = {
        "descriptor_dim": D,
        "GNN_layers": ["self", "cross"] * num_layers,
        "sinkhorn_iterations": iters,
        "match_threshold": 0.2,
    }
    model = ref_superglue.SuperGlue(cfg).eval()
    spec = spec_of(model)
    sd = syn.synth_state_dict(spec, seed, gain)
    if peaky:
        sd = syn.superglue_peaky_(sd, scale=peaky)
    model.load_state_dict(sd)
    desc0, desc1 = syn.synth_descriptor_pairs(seed + 1, B, M, N, D)
    with torch.no_grad():
        out = model(desc0.transpose(1, 2).contiguous(), desc1.transpose(1, 2).contiguous())
    np.savez_compressed(
        os.path.join(HERE, f"superglue_{name}.npz"),
        meta=json.dumps(dict(D=D, num_layers=num_layers, iters=iters, B=B, M=M, N=N, seed=seed, gain=gain, peaky=peaky, spec=spec)),
        P=out["P"].numpy(),
        matches0=out["matches0"].numpy(),
        matches1=out["matches1"].numpy(),
        matching_scores0=out["matching_scores0"].numpy(),
        matching_scores1=out["matching_scores1"].numpy(),
    )


def golden_language_encoder(ref_modules, name, D, n_queries, n_hints, seed):
    words = syn.known_words()
    model = ref_modules.LanguageEncoder(words, D, bi_dir=True).eval()
    spec = spec_of(model)
    model.load_state_dict(syn.synth_state_dict(spec, seed))
    texts = syn.synth_queries(seed + 1, n_queries, n_hints)
    texts[0] = texts[0] + " Some unknownword, here."  # OOV -> index 0, still consumes LSTM steps
    texts[1] = "The pose is north of a gray box."  # ragged: a short one
    with torch.no_grad():
        enc = model(texts)
    np.savez_compressed(
        os.path.join(HERE, f"language_encoder_{name}.npz"),
        meta=json.dumps(dict(D=D, seed=seed, spec=spec, texts=texts, words=words)),
        encodings=enc.numpy(),
    )


def golden_get_mlp(ref_modules, seed):
    out = {}
    metas = []
    for i, (channels, bn) in enumerate([([6, 32, 64], True), ([3, 64, 256], True), ([16, 8], False)]):
        model = ref_modules.get_mlp(channels, add_batchnorm=bn).eval()
        spec = spec_of(model)
        model.load_state_dict(syn.synth_state_dict(spec, seed + i))
        g = torch.Generator().manual_seed(seed + 100 + i)
        x = torch.randn(17, channels[0], generator=g)
        with torch.no_grad():
            out[f"y{i}"] = model(x).numpy()
        metas.append(dict(channels=channels, bn=bn, seed=seed + i, xseed=seed + 100 + i, spec=spec))
    np.savez_compressed(os.path.join(HERE, "get_mlp.npz"), meta=json.dumps(metas), **out)


def golden_retrieval(seed):
    """The reference's float64 numpy loop, verbatim (training/coarse.py:100-105,134-140)."""
    out = {}
    metas = []
    for i, (Q, N, D, k) in enumerate([(1, 128, 256, 10), (7, 1000, 256, 10), (64, 10000, 256, 10), (3, 37, 32, 5)]):
        cells = syn.synth_db_embeddings(seed + i, N, D).numpy()
        texts = syn.synth_query_embeddings(seed + 50 + i, Q, D).numpy()
        cell_encodings = np.zeros((N, D))
        text_encodings = np.zeros((Q, D))
        cell_encodings[:] = cells
        text_encodings[:] = texts
        top = np.zeros((Q, k), dtype=np.int64)
        for query_idx in range(len(text_encodings)):
            scores = cell_encodings[:] @ text_encodings[query_idx]
            sorted_indices = np.argsort(-1.0 * scores)  # High -> low
            top[query_idx] = sorted_indices[0:k]
        out[f"top{i}"] = top
        metas.append(dict(Q=Q, N=N, D=D, k=k, db_seed=seed + i, q_seed=seed + 50 + i))
    np.savez_compressed(os.path.join(HERE, "retrieval.npz"), meta=json.dumps(metas), **out)


def golden_pipeline(seed):
    """Run the reference's OWN ``evaluation/pipeline.py`` ``run_coarse`` / ``run_fine`` (imported unmodified, with the
    compat layer standing in for easydict / torch_geometric.transforms) over the CPU oracle models on a small synthetic scene;
    store retrievals, accuracy dicts and the matcher outputs.  ``training.coarse.eval_epoch`` (which evaluation.pipeline
    imports) resolves to this repository's CUDA drop-in, so the CPU restatement ``oracle.models.eval_epoch`` is patched in."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    import pipeline_common as pc
    from text2pos_cvpr2022_b200 import compat

    compat.install()
    if REF not in sys.path:
        sys.path.append(REF)  # after the repository: models.* / training.coarse resolve to the B200 modules
    import evaluation.pipeline as ref
    from datapreparation.kitti360pose.imports import Object3d

    ds, loader = pc.scene(seed)
    args = pc.pipeline_args()
    cm, csd = pc.coarse_state_dict()
    fm, fsd = pc.fine_state_dict()
    coarse = oracle.models.OracleCoarseModel(csd, cm.language_encoder.known_words)
    fine = pc.RecordingModel(oracle.models.OracleFineModel(fsd, fm.language_encoder.known_words))
    ref.eval_epoch_retrieval = oracle.models.eval_epoch
    retrievals, coarse_acc = ref.run_coarse(coarse, loader, args)
    ref.transform = pc.reference_transform()
    np.random.seed(seed)
    acc_mean, acc_off, acc_conf = ref.run_fine(fine, retrievals, loader, args)
    flat = lambda a: np.array([[float(a[k][t]) for t in sorted(a[k])] for k in sorted(a)])
    np.savez_compressed(
        os.path.join(HERE, "pipeline_small.npz"),
        meta=json.dumps(dict(seed=seed, n_cells=len(ds.all_cells), n_poses=len(ds.all_poses), top_k=args.top_k, threshs=args.threshs,
                             padding="Object3d.create_padding on np.random.seed(seed)", object3d=str(Object3d))),
        retrievals=np.array([list(r) for r in retrievals]), coarse_acc=flat(coarse_acc), acc_mean=flat(acc_mean),
        acc_offset=flat(acc_off), acc_mean_conf=flat(acc_conf),
        matches0=np.stack([o["matches0"] for o in fine.outputs]), offsets=np.stack([o["offsets"] for o in fine.outputs]),
        P=np.stack([o["P"] for o in fine.outputs]),
    )


def main():
    if "--pipeline-only" in sys.argv:
        golden_pipeline(seed=11)
        return
    ref_modules, ref_superglue = _import_reference()
    torch.manual_seed(0)
    golden_superglue(ref_superglue, "fine", D=128, num_layers=6, iters=50, B=8, M=16, N=6, seed=11, gain=0.4, peaky=5.0)
    golden_superglue(ref_superglue, "small", D=32, num_layers=2, iters=20, B=5, M=7, N=3, seed=12, gain=1.0, peaky=0)
    golden_language_encoder(ref_modules, "coarse", D=256, n_queries=12, n_hints=6, seed=21)
    golden_language_encoder(ref_modules, "fine", D=128, n_queries=6, n_hints=1, seed=22)
    golden_get_mlp(ref_modules, seed=31)
    golden_retrieval(seed=41)
    golden_pipeline(seed=11)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
