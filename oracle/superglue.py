"""Oracle (test infrastructure): the fine matcher's SuperGlue head.

Restates ``models/superglue.py`` of the reference in row-major ``[B, n, D]``
layout (the reference is channel-first ``[B, D, n]``):

* ``MultiHeadedAttention.forward`` :108-115 + ``attention`` :90-94 -- 4 heads;
  the channel -> (dim, head) split is ``view(B, D/4, 4, n)``, i.e. channel ``c``
  belongs to head ``c % 4`` at head-dim ``c // 4``; softmax over the source.
* ``AttentionalPropagation.forward`` :125-127 -- ``Conv1d(2D,2D) -> BN(eval) ->
  ReLU -> Conv1d(2D,D)`` on ``cat[x, message]``.
* ``AttentionalGNN.forward`` :138-146 -- both sides' deltas from the OLD
  descriptors, shared layer weights, names = ["self","cross"] * num_layers.
* ``SuperGlue.forward`` tail :280-330, ``log_optimal_transport`` :158-177,
  ``log_sinkhorn_iterations`` :149-155, mutual-nearest-neighbour matching with
  ``exp(max) > match_threshold`` (0.2).

PINNED against the reference module itself (``tests/golden/make_golden.py`` ->
``superglue_*.npz``).  Inference semantics: BatchNorm in eval mode (the
reference pipeline leaves the fine model in train mode -- a documented quirk
that is not reproduced, SURVEY.md "Parity quirk register").
"""
import math
from typing import Dict, List, Sequence

import torch

from .mlp import BN_EPS

NUM_HEADS = 4
MATCH_THRESHOLD = 0.2


def _conv1x1(sd, prefix, x):
    """Conv1d(kernel_size=1) on rows: x [B,n,Cin] -> [B,n,Cout]."""
    return x @ sd[prefix + "weight"].squeeze(-1).t() + sd[prefix + "bias"]


def multi_head_attention(sd, prefix, x, src):
    B, n, D = x.shape
    m = src.shape[1]
    dh = D // NUM_HEADS
    q = _conv1x1(sd, prefix + "proj.0.", x).reshape(B, n, dh, NUM_HEADS)
    k = _conv1x1(sd, prefix + "proj.1.", src).reshape(B, m, dh, NUM_HEADS)
    v = _conv1x1(sd, prefix + "proj.2.", src).reshape(B, m, dh, NUM_HEADS)
    scores = torch.einsum("bndh,bmdh->bhnm", q, k) / dh**0.5
    prob = torch.softmax(scores, dim=-1)
    out = torch.einsum("bhnm,bmdh->bndh", prob, v).reshape(B, n, D)
    return _conv1x1(sd, prefix + "merge.", out)


def attentional_propagation(sd, prefix, x, src):
    msg = multi_head_attention(sd, prefix + "attn.", x, src)
    h = _conv1x1(sd, prefix + "mlp.0.", torch.cat([x, msg], dim=-1))
    bn = prefix + "mlp.1."
    h = (h - sd[bn + "running_mean"]) / torch.sqrt(sd[bn + "running_var"] + BN_EPS) * sd[bn + "weight"] + sd[bn + "bias"]
    return _conv1x1(sd, prefix + "mlp.3.", torch.relu(h))


def attentional_gnn(sd, prefix, desc0, desc1, layer_names: Sequence[str]):
    for li, name in enumerate(layer_names):
        p = f"{prefix}layers.{li}."
        src0, src1 = (desc1, desc0) if name == "cross" else (desc0, desc1)
        d0 = attentional_propagation(sd, p, desc0, src0)
        d1 = attentional_propagation(sd, p, desc1, src1)
        desc0, desc1 = desc0 + d0, desc1 + d1
    return desc0, desc1


def log_optimal_transport(scores: torch.Tensor, alpha: torch.Tensor, iters: int) -> torch.Tensor:
    """scores [B,M,N] -> log-assignment [B,M+1,N+1] (dustbin row/col appended)."""
    B, M, N = scores.shape
    Z = torch.empty(B, M + 1, N + 1, dtype=scores.dtype)
    Z[:, :M, :N] = scores
    Z[:, M, :] = alpha
    Z[:, :, N] = alpha
    # the reference computes the marginals in float32 tensor arithmetic (:160-171)
    one = scores.new_tensor(1)
    ms, ns = M * one, N * one
    norm_t = -(ms + ns).log()
    log_mu = torch.cat([norm_t.expand(M), ns.log()[None] + norm_t])
    log_nu = torch.cat([norm_t.expand(N), ms.log()[None] + norm_t])
    u = torch.zeros(B, M + 1, dtype=scores.dtype)
    v = torch.zeros(B, N + 1, dtype=scores.dtype)
    for _ in range(iters):
        u = log_mu - torch.logsumexp(Z + v[:, None, :], dim=2)
        v = log_nu - torch.logsumexp(Z + u[:, :, None], dim=1)
    return Z + u[:, :, None] + v[:, None, :] - norm_t


def match(logP: torch.Tensor, threshold: float = MATCH_THRESHOLD):
    """Mutual nearest neighbours over the real rows/cols of the log-assignment (:312-322)."""
    s = logP[:, :-1, :-1]
    max0, max1 = s.max(2), s.max(1)
    i0, i1 = max0.indices, max1.indices
    B, M = i0.shape
    N = i1.shape[1]
    mutual0 = torch.arange(M)[None] == i1.gather(1, i0)
    mutual1 = torch.arange(N)[None] == i0.gather(1, i1)
    zero = s.new_tensor(0)
    ms0 = torch.where(mutual0, max0.values.exp(), zero)
    ms1 = torch.where(mutual1, ms0.gather(1, i1), zero)
    valid0 = mutual0 & (ms0 > threshold)
    valid1 = mutual1 & valid0.gather(1, i1)
    m0 = torch.where(valid0, i0, i0.new_tensor(-1))
    m1 = torch.where(valid1, i1, i1.new_tensor(-1))
    return m0, m1, ms0, ms1


def superglue_forward(
    sd: Dict[str, torch.Tensor],
    prefix: str,
    desc0: torch.Tensor,
    desc1: torch.Tensor,
    num_layers: int,
    sinkhorn_iters: int,
):
    """desc0 [B,M,D], desc1 [B,N,D] (row layout) -> dict(P, matches0, matches1, matching_scores0/1, scores)."""
    D = desc0.shape[-1]
    names = ["self", "cross"] * num_layers  # models/superglue_matcher.py:78
    d0, d1 = attentional_gnn(sd, prefix + "gnn.", desc0.float(), desc1.float(), names)
    md0 = _conv1x1(sd, prefix + "final_proj.", d0)
    md1 = _conv1x1(sd, prefix + "final_proj.", d1)
    scores = torch.einsum("bnd,bmd->bnm", md0, md1) / D**0.5
    logP = log_optimal_transport(scores, sd[prefix + "bin_score"], sinkhorn_iters)
    m0, m1, ms0, ms1 = match(logP)
    return dict(P=logP.exp(), matches0=m0, matches1=m1, matching_scores0=ms0, matching_scores1=ms1,
                scores=scores, desc0=d0, desc1=d1)


def superglue_match_forward(sd, hint_enc: torch.Tensor, obj_enc: torch.Tensor, num_layers: int, sinkhorn_iters: int):
    """Tail of ``SuperGlueMatch.forward`` (models/superglue_matcher.py:96-128) given the raw encodings.

    hint_enc [B,N,D] = LanguageEncoder outputs, obj_enc [B,M,D] = ObjectEncoder outputs (both un-normalised).
    """
    import torch.nn.functional as F

    h = F.normalize(hint_enc, dim=-1)
    o = F.normalize(obj_enc, dim=-1)
    out = superglue_forward(sd, "superglue.", o, h, num_layers, sinkhorn_iters)
    hid = torch.relu(h @ sd["mlp_offsets.0.weight"].t() + sd["mlp_offsets.0.bias"])
    out["offsets"] = hid @ sd["mlp_offsets.2.weight"].t() + sd["mlp_offsets.2.bias"]  # :74,117
    return out
