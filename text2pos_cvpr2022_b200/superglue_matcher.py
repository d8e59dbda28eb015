"""Drop-in for the reference's ``models/superglue_matcher.py`` (fine hints-to-objects matcher).

``SuperGlueMatch(known_classes, known_colors, known_words, args).forward(objects, hints, object_points)`` returns an
attribute dict with ``P, matches0, matches1, offsets, matching_scores0, matching_scores1`` as
``evaluation.pipeline.run_fine`` consumes them; ``get_pos_in_cell`` is the host helper imported by the pipeline.
Differences from the reference, by design: one batched LSTM call for all hint sentences (the reference loops per
sample), eval-mode BatchNorm (the reference pipeline leaves the model in train mode -- documented quirk).
"""
from typing import List

import numpy as np
import torch
import torch.nn as nn

from . import _lib, packing
from .modules import LanguageEncoder, lstm_encode, tokenize
from .object_encoder import ObjectEncoder, object_encoder_forward
from .runtime import AttrDict, PackedModule, arg
from .superglue import SuperGlue, superglue_forward
from .synthetic import PackedCells, pack_cells


def get_mlp_offset(dims: List[int], add_batchnorm=False) -> nn.Sequential:
    """Linear stack WITHOUT trailing ReLU (keys ``0``, ``2`` for [D, D/2, 2]); models/superglue_matcher.py:29-48."""
    layers = []
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            layers.append(nn.ReLU())
            if add_batchnorm:
                layers.append(nn.BatchNorm1d(dims[i + 1]))
    return nn.Sequential(*layers)


class SuperGlueMatch(PackedModule):
    def __init__(self, known_classes: List[str], known_colors: List[str], known_words: List[str], args):
        super().__init__()
        self.embed_dim = arg(args, "embed_dim")
        self.num_layers = arg(args, "num_layers", 6)
        self.sinkhorn_iters = arg(args, "sinkhorn_iters", 50)
        self.use_features = arg(args, "use_features", ["class", "color", "position"])
        self.args = args
        D = self.embed_dim
        self.object_encoder = ObjectEncoder(D, known_classes, known_colors, args)
        self.language_encoder = LanguageEncoder(known_words, D, bi_dir=True)
        self.mlp_offsets = get_mlp_offset([D, D // 2, 2])
        self.superglue = SuperGlue(
            {
                "descriptor_dim": D,
                "GNN_layers": ["self", "cross"] * self.num_layers,
                "sinkhorn_iterations": self.sinkhorn_iters,
                "match_threshold": 0.2,
            }
        )

    def _t2p_pack(self, sd):
        bb = packing.BlobBuilder()
        cfg = self.superglue.config
        desc = dict(
            pointnet=packing.pack_pointnet2(bb, sd, "object_encoder.pointnet.", self.object_encoder.pointnet.self_loop_quirk),
            objenc=packing.pack_object_encoder(bb, sd, "object_encoder.", self.embed_dim),
            lstm=packing.pack_lstm(bb, sd, "language_encoder."),
            superglue=packing.pack_superglue(bb, sd, "superglue.", list(cfg["GNN_layers"]), cfg["sinkhorn_iterations"], cfg["match_threshold"]),
            off1=bb.linear(packing._np64(sd["mlp_offsets.0.weight"]), packing._np64(sd["mlp_offsets.0.bias"])),
            off2=bb.linear(packing._np64(sd["mlp_offsets.2.weight"]), packing._np64(sd["mlp_offsets.2.bias"])),
        )
        return bb.finish(), desc

    def forward(self, objects, hints, object_points):
        batch_size, num_objects = len(objects), len(objects[0])
        if any(len(o) != num_objects for o in objects):
            raise ValueError("SuperGlueMatch: every sample must hold the same (padded) number of objects")
        num_hints = len(hints[0])
        if any(len(h) != num_hints for h in hints):
            raise ValueError("SuperGlueMatch: every sample must hold the same number of hints")
        dev = self.t2p_device()
        flat = [s for sample in hints for s in sample]
        tokens, lengths = tokenize(flat, self.language_encoder.known_words)
        tok = torch.from_numpy(tokens).pin_memory().to(dev, non_blocking=True)
        ln = torch.from_numpy(lengths).pin_memory().to(dev, non_blocking=True)
        cells = pack_cells(objects, object_points).to(dev)
        return self.forward_packed(cells, tok, ln, batch_size, num_objects, num_hints)

    def forward_packed(self, cells: PackedCells, tokens, lengths, batch_size, num_objects, num_hints):
        lib = _lib.load()
        weights, desc = self.t2p_packed()
        dev = cells.pos.device
        D = self.embed_dim
        # hints: LanguageEncoder + F.normalize (superglue_matcher.py:93-96), one batched call
        hint_enc = lstm_encode(weights, desc["lstm"], tokens, lengths, True, self).reshape(batch_size, num_hints, D)
        # objects: ObjectEncoder + reshape + F.normalize (:101-103)
        obj_enc = object_encoder_forward(weights, desc["pointnet"], desc["objenc"], cells, self)
        with torch.cuda.device(dev):
            _lib.check(lib.t2p_l2_normalize_rows(_lib.ptr(obj_enc), obj_enc.shape[0], D, D, _lib.stream_ptr(dev)), "l2_normalize_rows")
        obj_enc = obj_enc.reshape(batch_size, num_objects, D)
        out = superglue_forward(weights, desc["superglue"], obj_enc, hint_enc, self)
        # offsets = Linear -> ReLU -> Linear on the hint encodings (:74,117)
        flat = hint_enc.reshape(batch_size * num_hints, D)
        hid = torch.empty(flat.shape[0], desc["off1"].n, dtype=torch.float32, device=dev)
        off = torch.empty(flat.shape[0], 2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            s = _lib.stream_ptr(dev)
            _lib.check(lib.t2p_linear(weights.handle, desc["off1"], _lib.ptr(flat), flat.shape[0], D, 1, _lib.ptr(hid), hid.shape[1], s), "linear")
            _lib.check(lib.t2p_linear(weights.handle, desc["off2"], _lib.ptr(hid), hid.shape[0], hid.shape[1], 0, _lib.ptr(off), 2, s), "linear")
        outputs = AttrDict()
        outputs.P = out["P"]
        outputs.matches0 = out["matches0"]
        outputs.matches1 = out["matches1"]
        outputs.offsets = off.reshape(batch_size, num_hints, 2)
        outputs.matching_scores0 = out["matching_scores0"]
        outputs.matching_scores1 = out["matching_scores1"]
        return outputs

    @property
    def device(self):
        return next(self.mlp_offsets.parameters()).device

    def get_device(self):
        return next(self.mlp_offsets.parameters()).device


def get_pos_in_cell(objects, matches0, offsets):
    """Pose estimate in cell coordinates: mean over matched objects of (object centre + offset of its hint);
    (0.5, 0.5) without matches.  Host helper of ``evaluation.pipeline`` (models/superglue_matcher.py:138-161)."""
    preds = [
        objects[obj_idx].get_center()[0:2] + offsets[hint_idx]
        for obj_idx, hint_idx in enumerate(matches0)
        if hint_idx != -1
    ]
    return np.mean(preds, axis=0) if len(preds) > 0 else np.array((0.5, 0.5))
