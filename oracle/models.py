"""Oracle (test infrastructure): CPU modules with the reference's module API, backed by the oracle restatements.

``OracleCoarseModel`` / ``OracleFineModel`` expose ``encode_text`` / ``encode_objects`` / ``forward`` exactly as
``models/cell_retrieval.py:69-107`` and ``models/superglue_matcher.py:87-128`` do, computed with ``oracle.text`` /
``oracle.cells`` / ``oracle.superglue`` from a plain ``state_dict``.  They let the tests run the evaluation pipeline (the
reference's own ``evaluation/pipeline.py`` functions as well as this repository's restatement) end to end on the CPU and
compare it with the CUDA modules.  ``eval_epoch`` restates ``training/coarse.py:68-167`` with the verbatim float64 numpy loop.
PARITY: inherits the status of the functions it calls (PyG-dependent pieces unpinned, see ``oracle/__init__.py``).
"""
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

from . import cells as ocells
from . import superglue as osg
from . import text as otext


class _AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _pack(objects, object_points):
    """(List[List[Object3d]], List[Batch]) -> per-cell [n_o,P,3] rgb / pos tensors + centres / mean colours from the RAW points
    (models/object_encoder.py:121-131)."""
    rgb, pos = [], []
    for objs, b in zip(objects, object_points):
        n = len(objs)
        rgb.append(b.x.float().reshape(n, -1, 3))
        pos.append(b.pos.float().reshape(n, -1, 3))
    centers = torch.tensor(np.array([o.get_center() for c in objects for o in c]), dtype=torch.float)
    colors = torch.tensor(np.array([o.get_color_rgb() for c in objects for o in c]), dtype=torch.float)
    return rgb, pos, centers, colors


class OracleCoarseModel:
    def __init__(self, sd: Dict[str, torch.Tensor], known_words: Dict[str, int], embed_dim: int = 256):
        self.sd, self.known_words, self.embed_dim = sd, known_words, embed_dim

    def eval(self):
        return self

    @torch.no_grad()
    def encode_text(self, descriptions: List[str]) -> torch.Tensor:
        return otext.encode_text(self.sd, descriptions, self.known_words)

    @torch.no_grad()
    def encode_objects(self, objects, object_points) -> torch.Tensor:
        rgb, pos, centers, colors = _pack(objects, object_points)
        return ocells.encode_objects(self.sd, rgb, pos, centers, colors)


class OracleFineModel:
    def __init__(self, sd: Dict[str, torch.Tensor], known_words: Dict[str, int], embed_dim: int = 128, num_layers: int = 6,
                 sinkhorn_iters: int = 50):
        self.sd, self.known_words, self.embed_dim = sd, known_words, embed_dim
        self.num_layers, self.sinkhorn_iters = num_layers, sinkhorn_iters

    def eval(self):
        return self

    @torch.no_grad()
    def __call__(self, objects, hints, object_points):
        B, M = len(objects), len(objects[0])
        rgb, pos, centers, colors = _pack(objects, object_points)
        obj = ocells.object_encoder(self.sd, "object_encoder.", rgb, pos, centers, colors).reshape(B, M, self.embed_dim)
        hint = torch.stack([otext.language_encoder(self.sd, "language_encoder.", *otext.tokenize(h, self.known_words))
                            for h in hints])  # one LSTM call per sample, models/superglue_matcher.py:93-95
        out = osg.superglue_match_forward(self.sd, hint, obj, self.num_layers, self.sinkhorn_iters)
        return _AttrDict(P=out["P"], matches0=out["matches0"], matches1=out["matches1"], offsets=out["offsets"],
                         matching_scores0=out["matching_scores0"], matching_scores1=out["matching_scores1"])

    forward = __call__


@torch.no_grad()
def eval_epoch(model, dataloader, args, return_encodings: bool = False):
    """``training/coarse.py:68-167`` on the CPU: encodings into float64 holders, per-query mat-vec + argsort."""
    top_k = list(args.top_k)
    dataset = dataloader.dataset
    cells_dataset = dataset.get_cell_dataset()
    cells_dict = {cell.id: cell for cell in cells_dataset.cells}
    cell_size = cells_dataset.cells[0].cell_size
    text_enc, query_cell_ids = [], []
    for batch in dataloader:
        text_enc.append(model.encode_text(batch["texts"]).cpu().numpy())
        query_cell_ids.extend(batch["cell_ids"])
    text_encodings = np.zeros((len(query_cell_ids), model.embed_dim))
    text_encodings[:] = np.concatenate(text_enc)
    cell_enc, db_cell_ids = [], []
    bs = int(args.batch_size)
    for i0 in range(0, len(cells_dataset), bs):
        items = [cells_dataset[i] for i in range(i0, min(i0 + bs, len(cells_dataset)))]
        cell_enc.append(model.encode_objects([it["objects"] for it in items], [it["object_points"] for it in items]).cpu().numpy())
        db_cell_ids.extend(it["cell_ids"] for it in items)
    cell_encodings = np.zeros((len(db_cell_ids), model.embed_dim))
    cell_encodings[:] = np.concatenate(cell_enc)
    db_cell_ids = np.array(db_cell_ids, dtype="<U32")
    query_poses_w = np.array([pose.pose_w[0:2] for pose in dataset.all_poses])
    acc, acc_close, top = {k: [] for k in top_k}, {k: [] for k in top_k}, {}
    for q in range(len(text_encodings)):
        scores = cell_encodings[:] @ text_encodings[q]
        order = np.argsort(-1.0 * scores, kind="stable")[0:np.max(top_k)]
        ids = db_cell_ids[order]
        for k in top_k:
            acc[k].append(query_cell_ids[q] in ids[0:k])
        top[q] = ids
        d = np.linalg.norm(query_poses_w[q] - [cells_dict[c].get_center()[0:2] for c in ids], axis=1)
        for k in top_k:
            acc_close[k].append(np.any(d[0:k] <= cell_size / 2))
    acc = {k: np.mean(v) for k, v in acc.items()}
    acc_close = {k: np.mean(v) for k, v in acc_close.items()}
    if return_encodings:
        return acc, acc_close, top, cell_encodings, text_encodings
    return acc, acc_close, top


def get_pos_in_cell(objects, matches0, offsets):
    """``models/superglue_matcher.py:138-161`` verbatim in structure."""
    pos_in_cell_pred = []
    for obj_idx, hint_idx in enumerate(matches0):
        if hint_idx == -1:
            continue
        pos_in_cell_pred.append(objects[obj_idx].get_center()[0:2] + offsets[hint_idx])
    return np.mean(pos_in_cell_pred, axis=0) if len(pos_in_cell_pred) > 0 else np.array((0.5, 0.5))
