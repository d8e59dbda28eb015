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

    # ---- cached fine stage ----------------------------------------------------------------
    def encode_cells_packed(self, cells: PackedCells, num_objects: int) -> torch.Tensor:
        """PackedCells whose cells all hold ``num_objects`` (padded) objects -> ``[n_cells, num_objects, D]`` unit-norm object
        encodings: the query-independent half of ``forward`` (models/superglue_matcher.py:101-103)."""
        lib = _lib.load()
        weights, desc = self.t2p_packed()
        dev = cells.pos.device
        D = self.embed_dim
        obj_enc = object_encoder_forward(weights, desc["pointnet"], desc["objenc"], cells, self)
        with torch.cuda.device(dev):
            _lib.check(lib.t2p_l2_normalize_rows(_lib.ptr(obj_enc), obj_enc.shape[0], D, D, _lib.stream_ptr(dev)), "l2_normalize_rows")
        if obj_enc.shape[0] % num_objects != 0:
            raise ValueError("encode_cells: every cell must hold the same (padded) number of objects")
        return obj_enc.reshape(-1, num_objects, D)

    def encode_cells(self, objects, object_points):
        """(List[List[Object3d]], List[Batch]) of PADDED cells -> (encodings [n_cells, pad, D], centres [n_cells, pad, 2] float64
        = ``obj.get_center()[0:2]`` for the pose head)."""
        num_objects = len(objects[0])
        if any(len(o) != num_objects for o in objects):
            raise ValueError("encode_cells: every cell must hold the same (padded) number of objects")
        dev = self.t2p_device()
        enc = self.encode_cells_packed(pack_cells(objects, object_points).to(dev), num_objects)
        centers = np.array([[obj.get_center()[0:2] for obj in cell] for cell in objects], dtype=np.float64)
        return enc, torch.from_numpy(centers).to(dev)

    def forward_cached(self, cache, cell_idx: torch.Tensor, hints):
        """Fine stage over cached cell tables: ``cell_idx`` [Q, K] int64 (rows of ``cache.obj_enc`` / ``cache.centers``, on the
        device), ``hints`` = Q lists of hint strings.  The hint LSTM runs once per QUERY (the reference repeats it for each of the
        K retrieved cells, evaluation/pipeline.py:191 with eval.py:160), the SuperGlue head gathers its inputs from the tables,
        the offset MLP runs per query and the pose head (``get_pos_in_cell``) runs on the device.  Returns an attribute dict
        with the reference's keys for the B = Q*K samples (query-major) plus ``confidence, pos_mean, pos_offsets``."""
        lib = _lib.load()
        weights, desc = self.t2p_packed()
        dev = cache.obj_enc.device
        Q, K = cell_idx.shape
        if len(hints) != Q:
            raise ValueError("forward_cached: one list of hints per query")
        N = len(hints[0])
        if any(len(h) != N for h in hints):
            raise ValueError("forward_cached: every query must hold the same number of hints")
        M, D = cache.obj_enc.shape[1], self.embed_dim
        B = Q * K
        tokens, lengths = tokenize([s for h in hints for s in h], self.language_encoder.known_words)
        tok = torch.from_numpy(tokens).pin_memory().to(dev, non_blocking=True)
        ln = torch.from_numpy(lengths).pin_memory().to(dev, non_blocking=True)
        hint_enc = lstm_encode(weights, desc["lstm"], tok, ln, True, self)  # [Q*N, D] unit rows
        idx0 = cell_idx.reshape(-1).to(dev, torch.int64).contiguous()
        idx1 = torch.arange(Q, device=dev, dtype=torch.int64).repeat_interleave(K).contiguous()
        P = torch.empty(B, M + 1, N + 1, dtype=torch.float32, device=dev)
        m0 = torch.empty(B, M, dtype=torch.int64, device=dev)
        m1 = torch.empty(B, N, dtype=torch.int64, device=dev)
        s0 = torch.empty(B, M, dtype=torch.float32, device=dev)
        s1 = torch.empty(B, N, dtype=torch.float32, device=dev)
        hid = torch.empty(Q * N, desc["off1"].n, dtype=torch.float32, device=dev)
        off = torch.empty(Q * N, 2, dtype=torch.float32, device=dev)
        pos_mean = torch.empty(B, 2, dtype=torch.float64, device=dev)
        pos_off = torch.empty(B, 2, dtype=torch.float64, device=dev)
        conf = torch.empty(B, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            ws = self.t2p_workspace(lib.t2p_superglue_workspace(B, M, N, D), dev)
            _lib.check(lib.t2p_superglue_forward_gather(weights.handle, desc["superglue"], _lib.ptr(cache.obj_enc), _lib.ptr(idx0),
                                                        _lib.ptr(hint_enc), _lib.ptr(idx1), B, M, N, _lib.ptr(P), _lib.ptr(m0),
                                                        _lib.ptr(m1), _lib.ptr(s0), _lib.ptr(s1), None, _lib.ptr(ws), ws.numel(), st),
                       "superglue_forward_gather")
            _lib.check(lib.t2p_linear(weights.handle, desc["off1"], _lib.ptr(hint_enc), Q * N, D, 1, _lib.ptr(hid), hid.shape[1], st), "linear")
            _lib.check(lib.t2p_linear(weights.handle, desc["off2"], _lib.ptr(hid), Q * N, hid.shape[1], 0, _lib.ptr(off), 2, st), "linear")
            _lib.check(lib.t2p_pose_head(_lib.ptr(m0), B, M, N, _lib.ptr(off), _lib.ptr(idx1), _lib.ptr(cache.centers), _lib.ptr(idx0),
                                         _lib.ptr(pos_mean), _lib.ptr(pos_off), _lib.ptr(conf), st), "pose_head")
        out = AttrDict()
        out.P, out.matches0, out.matches1 = P, m0, m1
        out.offsets = off.reshape(Q, 1, N, 2).expand(Q, K, N, 2).reshape(B, N, 2)
        out.matching_scores0, out.matching_scores1 = s0, s1
        out.confidence, out.pos_mean, out.pos_offsets = conf, pos_mean, pos_off
        return out

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


def pose_head(matches0: torch.Tensor, offsets: torch.Tensor, centers: torch.Tensor):
    """Batched ``get_pos_in_cell`` on the device: matches0 [B,M] int64, offsets [B,N,2] float32, centers [B,M,2] float64 ->
    (pos_mean [B,2], pos_offsets [B,2] float64, confidence [B] int32)."""
    lib = _lib.load()
    _lib.require_cuda(matches0, "matches")
    dev = matches0.device
    B, M = matches0.shape
    N = offsets.shape[1]
    matches0 = matches0.to(torch.int64).contiguous()
    offsets = offsets.to(dev, torch.float32).contiguous()
    centers = centers.to(dev, torch.float64).contiguous()
    pos_mean = torch.empty(B, 2, dtype=torch.float64, device=dev)
    pos_off = torch.empty(B, 2, dtype=torch.float64, device=dev)
    conf = torch.empty(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.t2p_pose_head(_lib.ptr(matches0), B, M, N, _lib.ptr(offsets), None, _lib.ptr(centers), None,
                                     _lib.ptr(pos_mean), _lib.ptr(pos_off), _lib.ptr(conf), _lib.stream_ptr(dev)), "pose_head")
    return pos_mean, pos_off, conf


def pose_accuracy(cache, cell_idx: torch.Tensor, out, pose_w: torch.Tensor, pose_scene: torch.Tensor, top_k, threshs) -> torch.Tensor:
    """Batched ``calc_sample_accuracies`` (evaluation/utils.py:31-54) + the mean-conf variant on the device ->
    hits [3, Q, len(top_k), len(threshs)] int32 (0/1): in-cell mean, mean with offsets, mean of the most confident cell."""
    import ctypes as C

    lib = _lib.load()
    dev = cell_idx.device
    Q, K = cell_idx.shape
    nk, nt = len(top_k), len(threshs)
    hits = torch.empty(3, Q, nk, nt, dtype=torch.int32, device=dev)
    ck = (C.c_int32 * nk)(*[int(k) for k in top_k])
    ct = (C.c_double * nt)(*[float(t) for t in threshs])
    with torch.cuda.device(dev):
        _lib.check(lib.t2p_pose_accuracy(_lib.ptr(out.pos_mean), _lib.ptr(out.pos_offsets), _lib.ptr(out.confidence),
                                         _lib.ptr(cell_idx.to(torch.int64).contiguous()), Q, K, _lib.ptr(cache.origin),
                                         _lib.ptr(cache.cell_size), _lib.ptr(cache.scene), _lib.ptr(pose_w.contiguous()),
                                         _lib.ptr(pose_scene.contiguous()), ck, nk, ct, nt, _lib.ptr(hits), _lib.stream_ptr(dev)),
                   "pose_accuracy")
    return hits
