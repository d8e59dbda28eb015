"""Oracle (test infrastructure): object encoder + cell aggregation of the reference.

Restates ``ObjectEncoder.forward`` (``models/object_encoder.py:61-142``, default
flags: ``class_embed=color_embed=False``, ``use_features=[class,color,position]``,
``pointnet_features=2``) and ``CellRetrievalNetwork.encode_objects``
(``models/cell_retrieval.py:77-107``, ``variation == 0``).

PARITY UNPINNED for ``DynamicEdgeConv(k=8, aggr='max')`` / ``global_max_pool``
(torch_geometric; absent, unpinned).  Restated semantics: for every object the
``min(8, n_obj_in_cell)`` nearest objects of the SAME cell under squared
Euclidean distance of the (normalised) embeddings, self included; ties at the
k-th place -> lowest index (torch_cluster's insertion only replaces on strictly
smaller distance).  The distance is accumulated in float32 sequentially over the
channels with individually rounded ``diff*diff`` and ``+`` (no FMA).  message =
``nn(cat[x_i, x_j - x_i])``, aggregation = max, then per-cell max pool.
"""
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from .mlp import get_mlp
from .pointnet import pointnet2_features

KNN_K = 8  # models/cell_retrieval.py:46-48


def object_encoder(
    sd: Dict[str, torch.Tensor],
    prefix: str,
    cell_rgb: Sequence[torch.Tensor],
    cell_pos: Sequence[torch.Tensor],
    centers: torch.Tensor,
    mean_rgb: torch.Tensor,
    self_loop_quirk: bool = True,
) -> torch.Tensor:
    """cell_rgb/cell_pos: one [n_o,P,3] tensor per cell; centers/mean_rgb [sum n_o, 3] -> [sum n_o, D]."""
    feats = [
        pointnet2_features(sd, prefix + "pointnet.", rgb, pos, self_loop_quirk)
        for rgb, pos in zip(cell_rgb, cell_pos)
    ]  # one PointNet2 forward PER CELL, models/object_encoder.py:92-95
    f = torch.cat(feats, dim=0)
    f_pn = F.normalize(get_mlp(sd, prefix + "mlp_pointnet.", f), dim=-1)  # :98,111
    f_col = F.normalize(get_mlp(sd, prefix + "color_encoder.", mean_rgb.float()), dim=-1)  # :121-127
    f_pos = F.normalize(get_mlp(sd, prefix + "pos_encoder.", centers.float()), dim=-1)  # :129-135
    return get_mlp(sd, prefix + "mlp_merge.", torch.cat([f_pn, f_col, f_pos], dim=-1))  # :138


def seq_sqdist_f32(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """sum_c (a_c-b_c)^2 accumulated sequentially over c in float32, no FMA.  a,b [..., C]."""
    a = a.astype(np.float32, copy=False)
    b = b.astype(np.float32, copy=False)
    acc = np.zeros(np.broadcast_shapes(a.shape, b.shape)[:-1], dtype=np.float32)
    for c in range(a.shape[-1]):
        d = a[..., c] - b[..., c]
        acc = acc + d * d
    return acc


def knn_in_cell(e: np.ndarray, k: int = KNN_K) -> np.ndarray:
    """e [n, D] -> nbr [n, min(k,n)] int64: the k nearest (self included), ties -> lowest index."""
    d2 = seq_sqdist_f32(e[:, None, :], e[None, :, :])
    return np.argsort(d2, axis=1, kind="stable")[:, : min(k, e.shape[0])]


def cell_aggregate(
    sd: Dict[str, torch.Tensor], emb: torch.Tensor, cell_offsets: Sequence[int]
) -> torch.Tensor:
    """emb [sum n_o, D] (ObjectEncoder output, NOT yet normalised) -> cell embeddings [B, D]."""
    e = F.normalize(emb, dim=-1)  # models/cell_retrieval.py:94
    pooled = []
    for b in range(len(cell_offsets) - 1):
        ec = e[cell_offsets[b] : cell_offsets[b + 1]]
        nbr = torch.from_numpy(knn_in_cell(ec.numpy()))  # [n, k']
        xi = ec[:, None, :].expand(-1, nbr.shape[1], -1)
        xj = ec[nbr]
        msg = get_mlp(sd, "graph1.nn.", torch.cat([xi, xj - xi], dim=-1))  # :97
        x = msg.max(dim=1).values  # aggr='max'
        pooled.append(x.max(dim=0).values)  # global_max_pool :98
    x = get_mlp(sd, "lin.", torch.stack(pooled))  # :99 (BN + trailing ReLU)
    return F.normalize(x)  # :105


def encode_objects(
    sd, cell_rgb, cell_pos, centers, mean_rgb, self_loop_quirk: bool = True
) -> torch.Tensor:
    """``CellRetrievalNetwork.encode_objects`` on packed tensors (one [n_o,P,3] pair per cell)."""
    emb = object_encoder(sd, "object_encoder.", cell_rgb, cell_pos, centers, mean_rgb, self_loop_quirk)
    offsets = np.concatenate([[0], np.cumsum([int(r.shape[0]) for r in cell_rgb])]).tolist()
    return cell_aggregate(sd, emb, offsets)
