"""Drop-in for the reference's ``models/superglue.py`` (SuperGlue matching head).

``SuperGlue(config)`` keeps the reference's ``state_dict`` keys (``kenc.*`` -- constructed but never evaluated, as in
the reference --, ``gnn.layers.<L>.attn.{merge,proj.0,proj.1,proj.2}``, ``gnn.layers.<L>.mlp.{0,1,3}``, ``final_proj``,
``bin_score``) and ``forward(desc0 [B,D,M], desc1 [B,D,N]) -> dict``; all layers, the Sinkhorn iterations and the
matching run in ONE persistent kernel (``csrc/superglue.cu``).
"""
from copy import deepcopy

import torch
import torch.nn as nn

from . import _lib, packing
from .runtime import PackedModule


def _conv_mlp(channels, do_bn=True) -> nn.Sequential:
    """Conv1d(k=1) stack with BatchNorm+ReLU between layers (parameter container; indices 0,1,3 as upstream)."""
    layers = []
    last = len(channels) - 1
    for i in range(1, len(channels)):
        layers.append(nn.Conv1d(channels[i - 1], channels[i], kernel_size=1, bias=True))
        if i < last:
            if do_bn:
                layers.append(nn.BatchNorm1d(channels[i]))
            layers.append(nn.ReLU())
    return nn.Sequential(*layers)


class KeypointEncoder(nn.Module):
    def __init__(self, feature_dim, layers):
        super().__init__()
        self.encoder = _conv_mlp([3] + list(layers) + [feature_dim])
        nn.init.constant_(self.encoder[-1].bias, 0.0)


class MultiHeadedAttention(nn.Module):
    def __init__(self, num_heads: int, d_model: int):
        super().__init__()
        assert d_model % num_heads == 0
        self.dim = d_model // num_heads
        self.num_heads = num_heads
        self.merge = nn.Conv1d(d_model, d_model, kernel_size=1)
        self.proj = nn.ModuleList([deepcopy(self.merge) for _ in range(3)])


class AttentionalPropagation(nn.Module):
    def __init__(self, feature_dim: int, num_heads: int):
        super().__init__()
        self.attn = MultiHeadedAttention(num_heads, feature_dim)
        self.mlp = _conv_mlp([feature_dim * 2, feature_dim * 2, feature_dim])
        nn.init.constant_(self.mlp[-1].bias, 0.0)


class AttentionalGNN(nn.Module):
    def __init__(self, feature_dim: int, layer_names: list):
        super().__init__()
        self.layers = nn.ModuleList([AttentionalPropagation(feature_dim, 4) for _ in range(len(layer_names))])
        self.names = layer_names


class SuperGlue(PackedModule):
    default_config = {
        "descriptor_dim": 256,
        "weights": "indoor",
        "keypoint_encoder": [32, 64, 128, 256],
        "GNN_layers": ["self", "cross"] * 9,
        "sinkhorn_iterations": 100,
        "match_threshold": 0.2,
    }

    def __init__(self, config):
        super().__init__()
        self.config = {**self.default_config, **config}
        D = self.config["descriptor_dim"]
        self.kenc = KeypointEncoder(D, self.config["keypoint_encoder"])
        names = list(self.config["GNN_layers"])
        self.gnn = AttentionalGNN(D, names) if len(names) > 0 else None
        self.final_proj = nn.Conv1d(D, D, kernel_size=1, bias=True)
        self.register_parameter("bin_score", torch.nn.Parameter(torch.tensor(1.0)))

    def _t2p_pack(self, sd):
        bb = packing.BlobBuilder()
        names = list(self.config["GNN_layers"]) if self.gnn is not None else []
        desc = packing.pack_superglue(bb, sd, "", names, self.config["sinkhorn_iterations"], self.config["match_threshold"])
        return bb.finish(), desc

    def match_rows(self, desc0: torch.Tensor, desc1: torch.Tensor, return_scores: bool = False):
        """Row layout fast path: desc0 [B,M,D], desc1 [B,N,D] -> dict like ``forward``."""
        weights, desc = self.t2p_packed()
        return superglue_forward(weights, desc, desc0, desc1, self, return_scores)

    def forward(self, desc0, desc1):
        """desc0 [B,D,M], desc1 [B,D,N] channel-first as in the reference (models/superglue.py:239)."""
        return self.match_rows(desc0.transpose(1, 2).contiguous(), desc1.transpose(1, 2).contiguous())


def superglue_forward(weights, desc, desc0, desc1, owner: PackedModule, return_scores=False):
    lib = _lib.load()
    _lib.require_cuda(desc0, "descriptors")
    desc0 = desc0.float().contiguous()
    desc1 = desc1.float().contiguous()
    B, M, D = desc0.shape
    N = desc1.shape[1]
    if D != desc.dim or desc1.shape[2] != D or desc1.shape[0] != B:
        raise ValueError(f"SuperGlue: descriptor shapes {tuple(desc0.shape)} / {tuple(desc1.shape)} do not match dim {desc.dim}")
    dev = desc0.device
    P = torch.empty(B, M + 1, N + 1, dtype=torch.float32, device=dev)
    m0 = torch.empty(B, M, dtype=torch.int64, device=dev)
    m1 = torch.empty(B, N, dtype=torch.int64, device=dev)
    s0 = torch.empty(B, M, dtype=torch.float32, device=dev)
    s1 = torch.empty(B, N, dtype=torch.float32, device=dev)
    sc = torch.empty(B, M, N, dtype=torch.float32, device=dev) if return_scores else None
    with torch.cuda.device(dev):
        ws = owner.t2p_workspace(lib.t2p_superglue_workspace(B, M, N, D), dev)
        _lib.check(
            lib.t2p_superglue_forward(weights.handle, desc, _lib.ptr(desc0), _lib.ptr(desc1), B, M, N, _lib.ptr(P), _lib.ptr(m0),
                                      _lib.ptr(m1), _lib.ptr(s0), _lib.ptr(s1), _lib.ptr(sc), _lib.ptr(ws), ws.numel(),
                                      _lib.stream_ptr(dev)),
            "superglue_forward",
        )
    out = {"matches0": m0, "matches1": m1, "matching_scores0": s0, "matching_scores1": s1, "P": P}
    if return_scores:
        out["scores"] = sc
    return out
