"""Oracle (test infrastructure): PointNet++ object encoder of the reference.

Restates ``models/pointcloud/pointnet2.py`` of the reference
(``SetAbstractionLayer.forward`` :25-37, ``GlobalAbstractionLayer.forward``
:45-49, ``PointNet2.forward`` :80-100) for ONE PyG batch = the objects of one
cell (the reference calls PointNet2 once per cell:
``models/object_encoder.py:92-95``).

PARITY UNPINNED: ``fps``, ``radius`` and ``PointConv`` are torch_geometric /
torch_cluster ops (not in ``/root/reference``, no pinned version, not
installable here).  Their published semantics are restated with the tie-breaks
made explicit; every kernel of the CUDA path follows the SAME choices:

* ``fps(pos, batch, ratio=0.5)``: ``ceil(ratio*n)`` samples per object, in
  selection order.  Start index = 0 (the reference default is a random start);
  next = argmax over the running min of squared distances, ties -> LOWEST
  index.  Squared distance is evaluated in float32 as
  ``((dx*dx) + (dy*dy)) + (dz*dz)`` with every product/sum individually rounded
  (no FMA contraction).
* ``radius(x, y, r, max_num_neighbors=32)``: for each centre the first 32
  candidates of the SAME object in ascending index with ``d2 < r2`` (strict),
  ``r2 = float32(float64(r)*float64(r))``, same distance arithmetic as above
  (torch_cluster's CUDA kernel order; its CPU nanoflann path returns tree order
  and is NOT followed).
* ``PointConv(local_nn, add_self_loops=True)`` on the bipartite
  (all points -> centres) graph: edges whose source and target indices are
  numerically equal *in the flat per-batch index spaces* are removed, then an
  edge ``i -> i`` is added for every ``i < n_centres`` ("self-loop quirk": the
  source is flat point ``i`` of the batch, which for every object but the first
  belongs to ANOTHER object of the same cell).  message =
  ``local_nn(cat[x_j, pos_j - pos_i])``, aggregation = max.
  ``self_loop_quirk=False`` gives the ``add_self_loops=False`` behaviour.
* ``global_max_pool``: per-object max.
"""
import math
from typing import Dict, Tuple

import numpy as np
import torch

from .mlp import get_mlp, linear

MAX_NEIGHBORS = 32
SA_RATIO = 0.5
SA_RADII = (0.2, 0.3, 0.4)  # models/pointcloud/pointnet2.py:57-59


def radius_sq(r: float) -> np.float32:
    return np.float32(float(r) * float(r))


def sqdist_f32(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """((dx*dx)+(dy*dy))+(dz*dz) in float32, no fused multiply-add (numpy never fuses)."""
    a = a.astype(np.float32, copy=False)
    b = b.astype(np.float32, copy=False)
    dx = a[..., 0] - b[..., 0]
    dy = a[..., 1] - b[..., 1]
    dz = a[..., 2] - b[..., 2]
    d = dx * dx
    d = d + dy * dy
    d = d + dz * dz
    return d


def fps(pos: np.ndarray, m: int) -> np.ndarray:
    """Farthest point sampling, vectorised over objects.

    pos: [n_obj, P, 3] float32 -> idx [n_obj, m] int64 (selection order, idx[:,0] == 0).
    """
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    n, P, _ = pos.shape
    idx = np.zeros((n, m), dtype=np.int64)
    mind = np.full((n, P), np.inf, dtype=np.float32)
    ar = np.arange(n)
    last = pos[:, 0, :]
    for s in range(1, m):
        d = sqdist_f32(pos, last[:, None, :])
        mind = np.minimum(mind, d)
        nxt = np.argmax(mind, axis=1)  # first occurrence == lowest index on ties
        idx[:, s] = nxt
        last = pos[ar, nxt, :]
    return idx


def ball_query(pos: np.ndarray, cpos: np.ndarray, r: float, cap: int = MAX_NEIGHBORS):
    """pos [n,P,3], cpos [n,m,3] -> (nbr_idx [n,m,cap] int32 padded with -1, count [n,m] int32)."""
    d2 = sqdist_f32(pos[:, None, :, :], cpos[:, :, None, :])  # [n, m, P]
    mask = d2 < radius_sq(r)
    rank = np.cumsum(mask, axis=2)
    keep = mask & (rank <= cap)
    n, m, P = keep.shape
    count = keep.sum(axis=2).astype(np.int32)
    nbr = np.full((n, m, cap), -1, dtype=np.int32)
    o, c, p = np.nonzero(keep)
    slot = rank[o, c, p] - 1
    nbr[o, c, slot] = p
    return nbr, count


def sa_edges(nbr: np.ndarray, count: np.ndarray, P: int, self_loop_quirk: bool):
    """Flatten the neighbour table of ONE cell into edge lists.

    Returns (src_obj, src_pt, dst_obj, dst_ctr) int64 arrays.  With the quirk the
    edge set is: radius edges minus {flat src == flat dst} plus {flat i -> flat i}.
    """
    n, m, cap = nbr.shape
    o, c, s = np.nonzero(nbr >= 0)
    p = nbr[o, c, s].astype(np.int64)
    so, sp, do, dc = o.astype(np.int64), p, o.astype(np.int64), c.astype(np.int64)
    if self_loop_quirk:
        keep = (so * P + sp) != (do * m + dc)
        so, sp, do, dc = so[keep], sp[keep], do[keep], dc[keep]
        flat = np.arange(n * m, dtype=np.int64)  # centre flat index == source flat index
        so = np.concatenate([so, flat // P])
        sp = np.concatenate([sp, flat % P])
        do = np.concatenate([do, flat // m])
        dc = np.concatenate([dc, flat % m])
    return so, sp, do, dc


def set_abstraction(
    sd: Dict[str, torch.Tensor],
    prefix: str,
    x: torch.Tensor,
    pos: torch.Tensor,
    r: float,
    self_loop_quirk: bool = True,
    ratio: float = SA_RATIO,
):
    """One SetAbstractionLayer on the objects of one cell.

    x [n,P,C] f32, pos [n,P,3] f32 -> (x_out [n,m,C_out], pos_out [n,m,3], idx [n,m], nbr, count)
    """
    n, P, _ = pos.shape
    m = int(math.ceil(ratio * P))
    pos_np = pos.numpy()
    idx = fps(pos_np, m)
    cpos_np = np.take_along_axis(pos_np, idx[:, :, None], axis=1)
    nbr, count = ball_query(pos_np, cpos_np, r)
    so, sp, do, dc = sa_edges(nbr, count, P, self_loop_quirk)
    cpos = torch.from_numpy(cpos_np)
    so_t, sp_t, do_t, dc_t = (torch.from_numpy(a) for a in (so, sp, do, dc))
    feat = torch.cat([x[so_t, sp_t], pos[so_t, sp_t] - cpos[do_t, dc_t]], dim=1)
    msg = get_mlp(sd, prefix + "point_conv.local_nn.", feat)
    c_out = msg.shape[1]
    out = torch.full((n * m, c_out), -torch.inf, dtype=msg.dtype)
    flat = (do_t * m + dc_t)[:, None].expand(-1, c_out)
    out = out.scatter_reduce(0, flat, msg, reduce="amax", include_self=True)
    assert torch.isfinite(out).all(), "a centre without any edge"
    return out.reshape(n, m, c_out), cpos, idx, nbr, count


def global_abstraction(sd, prefix, x, pos):
    """cat(x,pos) -> get_mlp -> per-object max.  x [n,m,C], pos [n,m,3] -> [n, C_out]."""
    h = get_mlp(sd, prefix + "mlp.", torch.cat([x, pos], dim=2))
    return h.max(dim=1).values


def pointnet2_features(
    sd: Dict[str, torch.Tensor],
    prefix: str,
    rgb: torch.Tensor,
    pos: torch.Tensor,
    self_loop_quirk: bool = True,
    return_intermediates: bool = False,
):
    """``PointNet2.forward(...).features2`` for the objects of ONE cell.

    rgb [n,P,3], pos [n,P,3] float32 (post FixedPoints+NormalizeScale) -> [n, 256].
    The classifier heads (:91-92) are dead on this path and not evaluated.
    """
    inter = {}
    x, p = rgb.float(), pos.float()
    for li, r in enumerate(SA_RADII, start=1):
        x, p, idx, nbr, count = set_abstraction(sd, f"{prefix}sa{li}.", x, p, r, self_loop_quirk)
        if return_intermediates:
            inter[f"sa{li}"] = dict(x=x, pos=p, idx=idx, nbr=nbr, count=count)
    f0 = global_abstraction(sd, prefix + "ga.", x, p)
    f1 = torch.relu(linear(sd, prefix + "lin1.", f0))
    f2 = torch.relu(linear(sd, prefix + "lin2.", f1))
    if return_intermediates:
        inter.update(features0=f0, features1=f1, features2=f2)
        return f2, inter
    return f2
