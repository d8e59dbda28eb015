"""Drop-in for the reference's ``models/object_encoder.py``.

Same constructor, ``forward(objects, object_points)`` signature and ``state_dict`` keys.  Only the configuration the
hot path uses is implemented (``class_embed = color_embed = False``, ``use_features = [class, color, position]``,
``pointnet_features = 2``); other ablation flags raise ``NotImplementedError``.
"""
import os
from typing import List

import numpy as np
import torch
import torch.nn as nn

from . import _lib, packing
from .modules import get_mlp
from .pointnet2 import PointNet2, pointnet2_forward
from .runtime import PackedModule, arg
from .synthetic import COLOR_NAMES, PackedCells, pack_cells

MAX_OBJECTS_PER_CALL = 8192  # bounds the PointNet++ workspace (~200 KB per object)


def check_hot_path_args(args):
    feats = list(arg(args, "use_features", ["class", "color", "position"]))
    if sorted(feats) != ["class", "color", "position"] or arg(args, "class_embed", False) or arg(args, "color_embed", False):
        raise NotImplementedError(
            "B200 ObjectEncoder implements the reference's default feature set "
            "(use_features=[class,color,position], class_embed=color_embed=False); ablation flags are out of scope"
        )
    if arg(args, "pointnet_features", 2) != 2:
        raise NotImplementedError("B200 ObjectEncoder implements pointnet_features=2 (features2) only")


class ObjectEncoder(PackedModule):
    def __init__(self, embed_dim: int, known_classes: List[str], known_colors: List[str], args):
        super().__init__()
        check_hot_path_args(args)
        self.embed_dim = embed_dim
        self.args = args
        self.known_classes = {c: (i + 1) for i, c in enumerate(known_classes)}
        self.known_classes["<unk>"] = 0
        self.class_embedding = nn.Embedding(len(self.known_classes), embed_dim, padding_idx=0)
        self.known_colors = {c: i for i, c in enumerate(COLOR_NAMES)}
        self.known_colors["<unk>"] = 0
        self.color_embedding = nn.Embedding(len(self.known_colors), embed_dim, padding_idx=0)
        self.pos_encoder = get_mlp([3, 64, embed_dim])
        self.color_encoder = get_mlp([3, 64, embed_dim])
        self.pointnet = PointNet2(len(known_classes), len(known_colors), args)
        path = arg(args, "pointnet_path", None)
        if path:
            if not os.path.isfile(path):
                raise FileNotFoundError(f"pointnet_path {path!r} does not exist")
            self.pointnet.load_state_dict(torch.load(path, map_location="cpu"))
        if arg(args, "pointnet_freeze", False):
            self.pointnet.requires_grad_(False)
        self.mlp_pointnet = get_mlp([self.pointnet.dim2, embed_dim])
        self.mlp_merge = get_mlp([3 * embed_dim, embed_dim])

    def _t2p_pack(self, sd):
        bb = packing.BlobBuilder()
        pn = packing.pack_pointnet2(bb, sd, "pointnet.", self.pointnet.self_loop_quirk)
        oe = packing.pack_object_encoder(bb, sd, "", self.embed_dim)
        return bb.finish(), (pn, oe)

    def forward_packed(self, cells: PackedCells) -> torch.Tensor:
        """PackedCells (on the module's device) -> object embeddings [n_obj, D] (un-normalised, as the reference)."""
        weights, (pn, oe) = self.t2p_packed()
        return object_encoder_forward(weights, pn, oe, cells, self)

    def forward(self, objects, object_points):
        """objects: List[List[Object3d]], object_points: List[Batch] (one PyG batch per cell) -> [sum n_obj, D]."""
        cells = pack_cells(objects, object_points).to(self.t2p_device())
        return self.forward_packed(cells)

    @property
    def device(self):
        return next(self.class_embedding.parameters()).device

    def get_device(self):
        return next(self.class_embedding.parameters()).device


def obj_cell_start_from_offsets(cell_offsets: torch.Tensor) -> torch.Tensor:
    """[n_cells+1] int32 -> [n_obj] int32: index of the first object of each object's cell."""
    counts = (cell_offsets[1:] - cell_offsets[:-1]).long()
    return torch.repeat_interleave(cell_offsets[:-1], counts).to(torch.int32)


def obj_cell_start_host(offsets, c0: int, c1: int, device) -> torch.Tensor:
    """The same for cells [c0, c1) of HOST offsets, relative to the first object of the run; built on the host and copied
    (repeat_interleave on the device would synchronise to learn its output size)."""
    off = np.asarray(offsets[c0: c1 + 1], dtype=np.int64) - int(offsets[c0])
    start = np.repeat(off[:-1], np.diff(off)).astype(np.int32)
    return torch.from_numpy(start).pin_memory().to(device, non_blocking=True)


def cell_chunks(cell_offsets_host, max_objects: int = MAX_OBJECTS_PER_CALL):
    """Split cells into runs whose object count stays <= max_objects (the quirk couples objects of a cell)."""
    chunks, start = [], 0
    n_cells = len(cell_offsets_host) - 1
    for c in range(1, n_cells + 1):
        if cell_offsets_host[c] - cell_offsets_host[start] > max_objects and c - 1 > start:
            chunks.append((start, c - 1))
            start = c - 1
    chunks.append((start, n_cells))
    return chunks


def object_encoder_forward(weights, pn_desc, oe_desc, cells: PackedCells, owner: PackedModule) -> torch.Tensor:
    lib = _lib.load()
    _lib.require_cuda(cells.pos, "object points")
    dev = cells.pos.device
    n_obj = cells.pos.shape[0]
    D = oe_desc.embed_dim
    emb = torch.empty(n_obj, D, dtype=torch.float32, device=dev)
    off_host = cells.host_offsets()
    for c0, c1 in cell_chunks(off_host):
        o0, o1 = off_host[c0], off_host[c1]
        if o1 == o0:
            continue
        if o1 - o0 > 65535:
            raise RuntimeError("a single cell with more than 65535 objects is not supported")
        start = obj_cell_start_host(off_host, c0, c1, dev)
        f2 = pointnet2_forward(weights, pn_desc, cells.pos[o0:o1], cells.rgb[o0:o1], start, owner)
        n = o1 - o0
        with torch.cuda.device(dev):
            ws = owner.t2p_workspace(lib.t2p_object_embed_workspace(oe_desc, n), dev)
            _lib.check(
                lib.t2p_object_embed(weights.handle, oe_desc, _lib.ptr(f2), _lib.ptr(cells.centers[o0:o1].contiguous()),
                                     _lib.ptr(cells.mean_rgb[o0:o1].contiguous()), n, _lib.ptr(emb[o0:o1]), _lib.ptr(ws),
                                     ws.numel(), _lib.stream_ptr(dev)),
                "object_embed",
            )
    return emb
