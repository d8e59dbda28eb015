"""Drop-in for the reference's ``models/pointcloud/pointnet2.py`` (PointNet++ object encoder).

Same ``state_dict`` keys (``sa{1,2,3}.point_conv.local_nn.<i>.{0,1}.*``, ``ga.mlp.<i>.{0,1}.*``, ``lin1``, ``lin2``,
``class_classifier``, ``color_classifier``).  ``forward`` runs ``t2p_pointnet2_forward`` (fused FPS + ball query,
edge-GEMM set abstraction, global abstraction) instead of torch_geometric ops.
"""
import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from . import _lib, packing
from .modules import get_mlp
from .runtime import AttrDict, PackedModule, arg


class _PointConvParams(nn.Module):
    """Holds ``local_nn`` under the key torch_geometric's PointConv uses."""

    def __init__(self, mlp: nn.Sequential):
        super().__init__()
        self.local_nn = mlp


class SetAbstractionLayer(nn.Module):
    def __init__(self, ratio, radius, mlp):
        super().__init__()
        self.ratio, self.radius = ratio, radius
        self.point_conv = _PointConvParams(mlp)


class GlobalAbstractionLayer(nn.Module):
    def __init__(self, mlp):
        super().__init__()
        self.mlp = mlp


class PointNet2(PackedModule):
    def __init__(self, num_classes, num_colors, args):
        super().__init__()
        assert arg(args, "pointnet_layers", 3) == 3 and arg(args, "pointnet_variation", 0) == 0
        self.sa1 = SetAbstractionLayer(0.5, 0.2, get_mlp([3 + 3, 32, 64]))
        self.sa2 = SetAbstractionLayer(0.5, 0.3, get_mlp([64 + 3, 128, 128]))
        self.sa3 = SetAbstractionLayer(0.5, 0.4, get_mlp([128 + 3, 256, 256]))
        self.ga = GlobalAbstractionLayer(get_mlp([256 + 3, 512, 1024]))
        self.lin1 = nn.Linear(1024, 512)
        self.lin2 = nn.Linear(512, 256)
        self.class_classifier = nn.Linear(256, num_classes)  # dead on the hot path, kept for the state_dict
        self.color_classifier = nn.Linear(256, num_colors)
        self.dim0, self.dim1, self.dim2 = 1024, 512, 256
        # PointConv(add_self_loops=True) flat-index self loops of the reference's PyG version (oracle/pointnet.py)
        self.self_loop_quirk = bool(arg(args, "pointnet_self_loop_quirk", True))

    def _t2p_pack(self, sd):
        bb = packing.BlobBuilder()
        desc = packing.pack_pointnet2(bb, sd, "", self.self_loop_quirk)
        return bb.finish(), desc

    def features_packed(self, pos: torch.Tensor, rgb: torch.Tensor, obj_cell_start: torch.Tensor, debug: bool = False):
        weights, desc = self.t2p_packed()
        return pointnet2_forward(weights, desc, pos, rgb, obj_cell_start, self, debug)

    def forward(self, data):
        """``data``: PyG-Batch-like (``.x`` rgb, ``.pos``, ``.batch``) holding the objects of ONE cell, every object
        with the same number of points.  Returns an attribute dict with ``features2`` (the only output the
        encoders read; the classifier heads are not evaluated on this path)."""
        dev = self.device
        pos, rgb, batch = data.pos.to(dev), data.x.to(dev), data.batch.to(dev)
        n_obj = int(batch.max().item()) + 1 if batch.numel() else 0
        if n_obj == 0 or pos.shape[0] % n_obj:
            raise ValueError("PointNet2: every object must hold the same number of points (FixedPoints)")
        P = pos.shape[0] // n_obj
        start = torch.zeros(n_obj, dtype=torch.int32, device=dev)
        f2 = self.features_packed(pos.float().reshape(n_obj, P, 3), rgb.float().reshape(n_obj, P, 3), start)
        return AttrDict(features2=f2)

    @property
    def device(self):
        return next(self.lin1.parameters()).device


def pointnet2_forward(weights, desc, pos, rgb, obj_cell_start, owner: PackedModule, debug: bool = False):
    """pos/rgb [n_obj,P,3] float32 cuda, obj_cell_start [n_obj] int32 -> features2 [n_obj,256] (+ debug dict)."""
    lib = _lib.load()
    _lib.require_cuda(pos, "object points")
    pos, rgb = pos.contiguous(), rgb.contiguous()
    obj_cell_start = obj_cell_start.to(torch.int32).contiguous()
    n_obj, P, _ = pos.shape
    dev = pos.device
    out = torch.empty(n_obj, desc.lin2.n, dtype=torch.float32, device=dev)
    dbg = None
    null4 = (C.c_void_p * 3)(None, None, None)
    p_idx = p_nbr = p_cnt = p_x = null4
    if debug:
        dbg = {}
        m = P
        arrs = {"idx": [], "nbr": [], "cnt": [], "x": []}
        for l in range(3):
            m = (m + 1) // 2
            arrs["idx"].append(torch.empty(n_obj, m, dtype=torch.int32, device=dev))
            arrs["nbr"].append(torch.empty(n_obj, m, _lib.MAX_NEIGHBORS, dtype=torch.int32, device=dev))
            arrs["cnt"].append(torch.empty(n_obj, m, dtype=torch.int32, device=dev))
            arrs["x"].append(torch.empty(n_obj, m, desc.sa_l2[l].n, dtype=torch.float32, device=dev))
        p_idx = (C.c_void_p * 3)(*[t.data_ptr() for t in arrs["idx"]])
        p_nbr = (C.c_void_p * 3)(*[t.data_ptr() for t in arrs["nbr"]])
        p_cnt = (C.c_void_p * 3)(*[t.data_ptr() for t in arrs["cnt"]])
        p_x = (C.c_void_p * 3)(*[t.data_ptr() for t in arrs["x"]])
        dbg = arrs
    with torch.cuda.device(dev):
        ws = owner.t2p_workspace(lib.t2p_pointnet2_workspace(desc, n_obj, P), dev)
        _lib.check(
            lib.t2p_pointnet2_forward(weights.handle, desc, _lib.ptr(pos), _lib.ptr(rgb), _lib.ptr(obj_cell_start), n_obj, P,
                                      _lib.ptr(out), p_idx, p_nbr, p_cnt, p_x, _lib.ptr(ws), ws.numel(),
                                      _lib.stream_ptr(dev)),
            "pointnet2_forward",
        )
    return (out, dbg) if debug else out
