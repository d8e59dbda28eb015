"""Drop-in for the reference's ``models/cell_retrieval.py`` (coarse text-to-cell retrieval model).

``CellRetrievalNetwork(known_classes, known_colors, known_words, args)`` with ``encode_text`` / ``encode_objects`` /
``embed_dim`` / ``device`` / ``get_device()`` exactly as ``training.coarse.eval_epoch`` and
``evaluation.pipeline.run_coarse`` use them, and the reference's ``state_dict`` keys (``graph1.nn.*``, ``lin.*``,
``object_encoder.*``, ``language_encoder.*``).  All compute runs in the sm_100a library.
"""
from typing import List

import numpy as np
import torch
import torch.nn as nn

from . import _lib, packing
from .modules import LanguageEncoder, get_mlp, lstm_encode
from .object_encoder import ObjectEncoder, check_hot_path_args, object_encoder_forward
from .runtime import PackedModule, arg
from .synthetic import PackedCells, pack_cells


class _EdgeConvParams(nn.Module):
    """Holds the edge MLP under the key torch_geometric's DynamicEdgeConv uses (``graph1.nn``)."""

    def __init__(self, mlp: nn.Sequential, k: int):
        super().__init__()
        self.nn = mlp
        self.k = k


class CellRetrievalNetwork(PackedModule):
    def __init__(self, known_classes: List[str], known_colors: List[str], known_words: List[str], args):
        super().__init__()
        self.embed_dim = arg(args, "embed_dim")
        self.use_features = arg(args, "use_features", ["class", "color", "position"])
        self.variation = arg(args, "variation", 0)
        self.args = args
        if self.variation != 0:
            raise NotImplementedError("variation=1 (mean aggregation) is an ablation outside the hot path")
        check_hot_path_args(args)
        D = self.embed_dim
        self.graph1 = _EdgeConvParams(get_mlp([2 * D, D, D], add_batchnorm=True), k=_lib.KNN_K)
        self.lin = get_mlp([D, D, D])
        self.object_encoder = ObjectEncoder(D, known_classes, known_colors, args)
        self.language_encoder = LanguageEncoder(known_words, D, bi_dir=True)

    # ---- packing ---------------------------------------------------------------------------------
    def _t2p_pack(self, sd):
        bb = packing.BlobBuilder()
        desc = dict(
            pointnet=packing.pack_pointnet2(bb, sd, "object_encoder.pointnet.", self.object_encoder.pointnet.self_loop_quirk),
            objenc=packing.pack_object_encoder(bb, sd, "object_encoder.", self.embed_dim),
            cellagg=packing.pack_cell_aggregation(bb, sd, self.embed_dim),
            lstm=packing.pack_lstm(bb, sd, "language_encoder."),
        )
        return bb.finish(), desc

    # ---- text side -------------------------------------------------------------------------------
    def encode_text(self, descriptions) -> torch.Tensor:
        """List[str] -> [B, D] unit-norm rows (models/cell_retrieval.py:69-75)."""
        from .modules import tokenize

        tokens, lengths = tokenize(descriptions, self.language_encoder.known_words)
        if len(lengths) and int(lengths.min()) < 1:
            raise ValueError("empty description (the reference's packed LSTM rejects length 0 too)")
        dev = self.t2p_device()
        tok = torch.from_numpy(tokens).pin_memory().to(dev, non_blocking=True)
        ln = torch.from_numpy(lengths).pin_memory().to(dev, non_blocking=True)
        return self.encode_tokens(tok, ln)

    def encode_tokens(self, tokens: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
        """Fast path: int32 tokens [B,T] / lengths [B] already on the device -> [B, D] unit-norm rows."""
        weights, desc = self.t2p_packed()
        return lstm_encode(weights, desc["lstm"], tokens, lengths, True, self)

    # ---- cell side -------------------------------------------------------------------------------
    def encode_objects(self, objects, object_points) -> torch.Tensor:
        """(List[List[Object3d]], List[Batch]) -> [B, D] unit-norm, non-negative rows (models/cell_retrieval.py:77-107)."""
        cells = pack_cells(objects, object_points).to(self.t2p_device())
        return self.encode_cells_packed(cells)

    def encode_cells_packed(self, cells: PackedCells, return_debug: bool = False):
        weights, desc = self.t2p_packed()
        emb = object_encoder_forward(weights, desc["pointnet"], desc["objenc"], cells, self)
        return cell_aggregate(weights, desc["cellagg"], emb, cells.cell_offsets, self, return_debug, cells.host_offsets())

    def forward(self):
        raise Exception("Not implemented.")

    @property
    def device(self):
        return next(self.lin.parameters()).device

    def get_device(self):
        return next(self.lin.parameters()).device


def cell_aggregate(weights, desc, emb: torch.Tensor, cell_offsets: torch.Tensor, owner: PackedModule, return_debug=False,
                   offsets_host=None):
    """``offsets_host``: the same offsets as a host list (saves the device sync of reading ``cell_offsets`` back)."""
    lib = _lib.load()
    _lib.require_cuda(emb, "object embeddings")
    dev = emb.device
    n_obj, D = emb.shape
    off_host = np.asarray(cell_offsets.tolist() if offsets_host is None else offsets_host, dtype=np.int64)
    n_cells = off_host.size - 1
    counts = np.diff(off_host)
    max_obj = int(counts.max()) if n_cells > 0 else 0
    if n_cells > 0 and int(counts.min()) < 1:
        raise ValueError("every cell needs at least one object")
    off_dev = cell_offsets.to(dev, torch.int32).contiguous()
    out = torch.empty(n_cells, D, dtype=torch.float32, device=dev)
    knn = torch.empty(n_obj, _lib.KNN_K, dtype=torch.int32, device=dev) if return_debug else None
    with torch.cuda.device(dev):
        ws = owner.t2p_workspace(lib.t2p_cell_aggregate_workspace(desc, n_obj, n_cells), dev)
        _lib.check(
            lib.t2p_cell_aggregate(weights.handle, desc, _lib.ptr(emb.contiguous()), _lib.ptr(off_dev), n_obj, n_cells, max_obj,
                                   _lib.ptr(out), _lib.ptr(knn), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)),
            "cell_aggregate",
        )
    return (out, dict(knn=knn, emb=emb)) if return_debug else out
