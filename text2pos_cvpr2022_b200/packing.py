"""state_dict -> packed float32 blob + descriptor structs for the C ABI.

All folding is done once, on the CPU, in float64 and rounded to float32 at the end:
* eval-mode BatchNorm1d folded into the preceding Linear / Conv1d(k=1)
  (``y = (xW^T + b - mean) * gamma / sqrt(var + eps) + beta``, eps = 1e-5);
* weights transposed to ``[K, N]`` row-major so that kernel B-tile loads are coalesced;
* LSTM input projection folded per vocabulary entry:
  ``xproj[dir, v] = emb[v] @ W_ih^T + b_ih + b_hh`` (gate order i,f,g,o);
* DynamicEdgeConv first layer split: ``W1 [x_i, x_j - x_i] = (W1a - W1b) x_i + W1b x_j``.

The key layout is the reference's (SURVEY.md 8a "state_dict layout").
"""
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib

BN_EPS = 1e-5
SA_RADII = (0.2, 0.3, 0.4)  # models/pointcloud/pointnet2.py:57-59 of the reference


def _np64(t: torch.Tensor) -> np.ndarray:
    return t.detach().to("cpu", torch.float64).numpy()


class BlobBuilder:
    ALIGN = 64  # floats (256 bytes)

    def __init__(self):
        self.chunks = []
        self.n = 0

    def add(self, arr: np.ndarray) -> int:
        """float64 values, rounded to float32 once, here."""
        return self._append(np.ascontiguousarray(arr, dtype=np.float64).reshape(-1).astype(np.float32))

    def add_raw_u32(self, arr: np.ndarray) -> int:
        """raw 32-bit words (e.g. packed fp16 pairs) stored bit-exactly inside the float32 blob."""
        a = np.ascontiguousarray(arr).reshape(-1)
        assert a.dtype == np.uint32
        return self._append(a.view(np.float32))

    def _append(self, a: np.ndarray) -> int:
        off = self.n
        pad = (-a.size) % self.ALIGN
        self.chunks.append(a)
        if pad:
            self.chunks.append(np.zeros(pad, dtype=np.float32))
        self.n += a.size + pad
        return off

    def linear(self, W: np.ndarray, b: Optional[np.ndarray], bn: Optional[Dict[str, np.ndarray]] = None) -> _lib.LinearDesc:
        """W [out, in] (nn.Linear layout), optional BN (dict weight,bias,running_mean,running_var)."""
        W = np.asarray(W, dtype=np.float64)
        out_f, in_f = W.shape
        b = np.zeros(out_f) if b is None else np.asarray(b, dtype=np.float64)
        if bn is not None:
            s = bn["weight"] / np.sqrt(bn["running_var"] + BN_EPS)
            W = W * s[:, None]
            b = (b - bn["running_mean"]) * s + bn["bias"]
        d = _lib.LinearDesc()
        d.w_off = self.add(W.T)
        d.b_off = self.add(b)
        d.k, d.n = in_f, out_f
        return d

    def finish(self) -> torch.Tensor:
        blob = np.concatenate(self.chunks) if self.chunks else np.zeros(1, dtype=np.float32)
        return torch.from_numpy(blob)


def _bn(sd, prefix) -> Optional[Dict[str, np.ndarray]]:
    if prefix + "running_mean" not in sd:
        return None
    return {k: _np64(sd[prefix + k]) for k in ("weight", "bias", "running_mean", "running_var")}


def pack_mlp_layer(bb: BlobBuilder, sd, prefix: str) -> _lib.LinearDesc:
    """One ``Sequential(Linear, [BatchNorm1d], ReLU)`` block of ``get_mlp`` stored under ``prefix`` (e.g. "lin.0.")."""
    return bb.linear(_np64(sd[prefix + "0.weight"]), _np64(sd[prefix + "0.bias"]), _bn(sd, prefix + "1."))


def pack_pointnet2(bb: BlobBuilder, sd, prefix: str, self_loop_quirk: bool = True) -> _lib.PointNet2Desc:
    d = _lib.PointNet2Desc()
    for l in range(3):
        p = f"{prefix}sa{l + 1}.point_conv.local_nn."
        d.sa_l1[l] = pack_mlp_layer(bb, sd, p + "0.")
        d.sa_l2[l] = pack_mlp_layer(bb, sd, p + "1.")
        d.sa_radius_sq[l] = float(np.float32(float(SA_RADII[l]) * float(SA_RADII[l])))
    d.ga_l1 = pack_mlp_layer(bb, sd, prefix + "ga.mlp.0.")
    d.ga_l2 = pack_mlp_layer(bb, sd, prefix + "ga.mlp.1.")
    d.lin1 = bb.linear(_np64(sd[prefix + "lin1.weight"]), _np64(sd[prefix + "lin1.bias"]))
    d.lin2 = bb.linear(_np64(sd[prefix + "lin2.weight"]), _np64(sd[prefix + "lin2.bias"]))
    d.self_loop_quirk = 1 if self_loop_quirk else 0
    for l in range(3):
        k, n = d.sa_l2[l].k, d.sa_l2[l].n
        d.sa_l2_tc_off[l] = -1
        if (k, n) in ((32, 64), (128, 128), (256, 256)):
            blob_w = np.concatenate(bb.chunks)[d.sa_l2[l].w_off: d.sa_l2[l].w_off + k * n].astype(np.float64).reshape(k, n)
            if fits_fp16_split(blob_w, SA_TC_WSCALE):  # else: the exact-fp32 kernel serves this layer
                if k == 32:      # K padded to one 64-wide chunk, the column block to 128 output channels (zeros)
                    pad = np.zeros((64, 128))
                    pad[:32, :64] = blob_w
                    img = _sa_tc_images(pad)
                elif k == 256:   # two 128-column blocks (one CTA each), every block with its four K chunks
                    img = np.concatenate([_sa_tc_images(blob_w[:, j: j + 128]) for j in (0, 128)])
                else:
                    img = _sa_tc_images(blob_w)
                d.sa_l2_tc_off[l] = bb.add_raw_u32(img)
    k, n = d.ga_l2.k, d.ga_l2.n
    d.ga_l2_tc_off = -1
    if k == 512 and n % 128 == 0:
        blob_w = np.concatenate(bb.chunks)[d.ga_l2.w_off: d.ga_l2.w_off + k * n].astype(np.float64).reshape(k, n)
        if fits_fp16_split(blob_w, SA_TC_WSCALE):
            d.ga_l2_tc_off = bb.add_raw_u32(np.concatenate([_sa_tc_images(blob_w[:, j: j + 128]) for j in range(0, n, 128)]))
    # x part of the dense layers (the last 3 rows of sa / ga first layers multiply pos and stay fp32): (layer, K of the x part, N)
    dense = [(d.sa_l1[1], 64, 128), (d.sa_l1[2], 128, 256), (d.ga_l1, 256, 512), (d.lin1, 1024, 512), (d.lin2, 512, 256)]
    for i in range(6):
        d.dense_tc_off[i] = -1
    flat = np.concatenate(bb.chunks)
    for i, (lin, kx, n) in enumerate(dense):
        if lin.n != n or lin.k not in (kx, kx + 3):
            continue
        w = flat[lin.w_off: lin.w_off + lin.k * lin.n].astype(np.float64).reshape(lin.k, lin.n)[:kx]
        if fits_fp16_split(w, SA_TC_WSCALE):
            d.dense_tc_off[i] = bb.add_raw_u32(np.concatenate([_sa_tc_images(w[:, j: j + 128]) for j in range(0, n, 128)]))
    return d


FP16_SAFE_MAX = 60000.0  # fp16 max is 65504


def fits_fp16_split(w: np.ndarray, scale: float) -> bool:
    """True if ``scale * w`` can be split into fp16 hi + lo images: finite and |scale * w| < 60000 (BatchNorm folding can
    blow weights up: a tiny running_var gives factors up to ~316 * gamma).  Layers that do not fit keep their tensor-core
    offset at -1 and run on the exact-fp32 kernels."""
    w = np.asarray(w, dtype=np.float64)
    return bool(np.isfinite(w).all() and (np.abs(w) * scale).max(initial=0.0) < FP16_SAFE_MAX)


SA_TC_WSCALE = 256.0  # 2^8: keeps the fp16 "lo" halves of the weights out of the subnormal range


def _sa_tc_images(w_kn: np.ndarray) -> np.ndarray:
    """w_kn [C, C] (= the BN-folded second local_nn layer as stored in the blob, [K, N]) -> uint32 words of the UMMA B-operand
    images streamed by ``sa_edge_tc_kernel``: [K chunk C/64][hi|lo][row n (output channel) C][64 fp16], value = fp16 split of
    2^8 * W[k = 64*chunk + e][n], the eight 16-byte units of every 128-byte row stored at (unit XOR (n & 7))."""
    K, N = w_kn.shape  # the set-abstraction layers are square; the global-abstraction block is [512, 256]
    assert K % 64 == 0 and N % 8 == 0
    w = w_kn.T * SA_TC_WSCALE                      # [n, k]
    hi = w.astype(np.float16)
    lo = (w - hi.astype(np.float64)).astype(np.float16)
    n = np.arange(N)
    swz = np.arange(8)[None, :] ^ (n[:, None] & 7)  # physical unit v of row n holds logical unit v ^ (n & 7)
    out = np.zeros((K // 64, 2, N, 8, 8), dtype=np.float16)
    for part, mat in enumerate((hi, lo)):
        t = mat.reshape(N, K // 64, 8, 8).transpose(1, 0, 2, 3)  # [chunk, n, logical unit, elem]
        out[:, part] = np.take_along_axis(t, swz[None, :, :, None], axis=2)
    return out.reshape(-1).view(np.uint32)


def pack_object_encoder(bb: BlobBuilder, sd, prefix: str, embed_dim: int) -> _lib.ObjEncDesc:
    d = _lib.ObjEncDesc()
    d.mlp_pointnet = pack_mlp_layer(bb, sd, prefix + "mlp_pointnet.0.")
    d.color_l1 = pack_mlp_layer(bb, sd, prefix + "color_encoder.0.")
    d.color_l2 = pack_mlp_layer(bb, sd, prefix + "color_encoder.1.")
    d.pos_l1 = pack_mlp_layer(bb, sd, prefix + "pos_encoder.0.")
    d.pos_l2 = pack_mlp_layer(bb, sd, prefix + "pos_encoder.1.")
    d.merge = pack_mlp_layer(bb, sd, prefix + "mlp_merge.0.")
    d.embed_dim = embed_dim
    return d


def pack_cell_aggregation(bb: BlobBuilder, sd, embed_dim: int) -> _lib.CellAggDesc:
    D = embed_dim
    d = _lib.CellAggDesc()
    W1 = _np64(sd["graph1.nn.0.0.weight"])  # [D, 2D] acting on cat[x_i, x_j - x_i]
    b1 = _np64(sd["graph1.nn.0.0.bias"])
    bn = _bn(sd, "graph1.nn.0.1.")
    if bn is not None:
        s = bn["weight"] / np.sqrt(bn["running_var"] + BN_EPS)
        W1 = W1 * s[:, None]
        b1 = (b1 - bn["running_mean"]) * s + bn["bias"]
    W1a, W1b = W1[:, :D], W1[:, D:]
    Wab = np.concatenate([W1a - W1b, W1b], axis=0)  # [2D, D] "nn.Linear layout": rows = outputs [A | B]
    d.edge_ab = bb.linear(Wab, np.concatenate([b1, np.zeros(D)]))
    d.edge_l2 = pack_mlp_layer(bb, sd, "graph1.nn.1.")
    d.lin_l1 = pack_mlp_layer(bb, sd, "lin.0.")
    d.lin_l2 = pack_mlp_layer(bb, sd, "lin.1.")
    d.embed_dim = D
    return d


def pack_lstm(bb: BlobBuilder, sd, prefix: str) -> _lib.LstmDesc:
    emb = _np64(sd[prefix + "word_embedding.weight"])  # [V, D]
    V, D = emb.shape
    xproj, whh = [], []
    for suffix in ("", "_reverse"):
        w_ih = _np64(sd[f"{prefix}lstm.weight_ih_l0{suffix}"])  # [4H, D]
        w_hh = _np64(sd[f"{prefix}lstm.weight_hh_l0{suffix}"])  # [4H, H]
        b = _np64(sd[f"{prefix}lstm.bias_ih_l0{suffix}"]) + _np64(sd[f"{prefix}lstm.bias_hh_l0{suffix}"])
        xproj.append(emb @ w_ih.T + b)  # [V, 4H]
        whh.append(w_hh.T)  # [H, 4H]
    H = whh[0].shape[0]
    d = _lib.LstmDesc()
    d.xproj_off = bb.add(np.stack(xproj))
    d.whh_off = bb.add(np.stack(whh))
    d.whh_reg_off = bb.add(_lstm_register_tiling(np.stack(whh))) if H in (32, 64, 128, 256) else -1
    if H == 256 and fits_fp16_split(np.stack(whh), LSTM_TC_WSCALE):  # else: the register-resident fp32 kernel
        xp = np.stack(xproj)  # [2, V, 4H], columns g*H + u
        d.xproj4_off = bb.add(xp.reshape(2, V, 4, H).transpose(0, 1, 3, 2))  # [2, V, H, 4]
        d.whh_tc_off = bb.add_raw_u32(_lstm_tc_images(np.stack(whh)))
    else:
        d.xproj4_off = d.whh_tc_off = -1
    d.vocab, d.hidden = V, H
    d.path = 0
    d.max_groups = 0
    return d


LSTM_TC_WSCALE = 256.0  # 2^8: keeps the fp16 "lo" halves of the weights out of the subnormal range


def _lstm_tc_images(whh_t: np.ndarray) -> np.ndarray:
    """whh_t [2, 256, 1024] (= W_hh^T) -> uint32 words of the tensor-memory images loaded by ``lstm_tc_kernel``:
    [dir 2][cluster rank 8][hi|lo 2][k-unit 32][row m 128][8 fp16], row m = 4*unit + gate of the rank's 32 hidden units,
    value = fp16 split of 2^8 * W_hh[gate*H + 32 r + unit][8*ku + e].  A warp reads 32 consecutive rows of one k-unit
    (512 contiguous bytes); the 8 fp16 of a row are 4 TMEM columns (element k in the (k%2) half of column k/2)."""
    H = whh_t.shape[1]
    assert H == 256
    m = np.arange(128)
    unit, gate = m // 4, m % 4
    out = np.zeros((2, 8, 2, 32, 128, 8), dtype=np.float16)
    for d in range(2):
        for r in range(8):
            cols = gate * H + 32 * r + unit                    # [128]
            w = whh_t[d][:, cols].T * LSTM_TC_WSCALE            # [128 rows, 256 k]
            hi = w.astype(np.float16)
            lo = (w - hi.astype(np.float64)).astype(np.float16)
            for part, mat in enumerate((hi, lo)):
                out[d, r, part] = mat.reshape(128, 32, 8).transpose(1, 0, 2)
    return out.reshape(-1).view(np.uint32)


def _lstm_register_tiling(whh_t: np.ndarray) -> np.ndarray:
    """whh_t [2, H, 4H] (= W_hh^T) -> [2, CS, NI, 4, 256, 4], the order in which the threads of ``lstm_reg_kernel``
    (csrc/lstm.cu) load their register-resident slice with coalesced float4 reads: cluster rank r owns hidden units
    [32r, 32r+32); thread w*32 + kp*4 + jj owns unit 32r + 4w + jj and the k values 4*(8i + kp) + e."""
    H = whh_t.shape[1]
    CS, NI = H // 32, H // 32
    r = np.arange(CS)[:, None, None, None, None]
    i = np.arange(NI)[None, :, None, None, None]
    e = np.arange(4)[None, None, :, None, None]
    t = np.arange(256)[None, None, None, :, None]
    g = np.arange(4)[None, None, None, None, :]
    w, kp, jj = t // 32, (t % 32) // 4, t % 4
    k = 4 * (8 * i + kp) + e
    col = g * H + 32 * r + 4 * w + jj
    k, col = np.broadcast_arrays(k, col)
    return np.stack([whh_t[d][k, col] for d in range(2)])


def pack_superglue(bb: BlobBuilder, sd, prefix: str, layer_names, sinkhorn_iters: int, match_threshold: float) -> _lib.SuperGlueDesc:
    d = _lib.SuperGlueDesc()
    n_layers = len(layer_names)
    if n_layers > _lib.MAX_GNN_LAYERS:
        raise ValueError(f"SuperGlue: {n_layers} GNN layers > {_lib.MAX_GNN_LAYERS}")

    def conv(p, bn=None):
        return bb.linear(_np64(sd[p + "weight"]).squeeze(-1), _np64(sd[p + "bias"]), bn)

    for L, name in enumerate(layer_names):
        p = f"{prefix}gnn.layers.{L}."
        d.q[L] = conv(p + "attn.proj.0.")
        d.k[L] = conv(p + "attn.proj.1.")
        d.v[L] = conv(p + "attn.proj.2.")
        d.merge[L] = conv(p + "attn.merge.")
        d.mlp0[L] = conv(p + "mlp.0.", _bn(sd, p + "mlp.1."))
        d.mlp3[L] = conv(p + "mlp.3.")
        d.is_cross[L] = 1 if name == "cross" else 0
    d.final_proj = conv(prefix + "final_proj.")
    d.num_gnn_layers = n_layers
    d.dim = int(sd[prefix + "final_proj.weight"].shape[0])
    d.sinkhorn_iters = int(sinkhorn_iters)
    d.bin_score = float(sd[prefix + "bin_score"])
    d.match_threshold = float(match_threshold)
    d.tc_w_off, d.tc_b_off = _pack_superglue_tc(bb, d, n_layers)
    return d


SG_TC_WSCALE = 256.0


def _pack_superglue_tc(bb: BlobBuilder, d: "_lib.SuperGlueDesc", n_layers: int):
    """Tensor-core weight stream + bias block of ``superglue_tc_kernel`` (csrc/superglue_tc.cu) from the BN-folded [K, N]
    matrices already in the blob; (-1, -1) unless D == 128 and every matrix fits the fp16 hi/lo split."""
    D = d.dim
    if D != 128:
        return -1, -1
    blob = np.concatenate(bb.chunks)

    def mat(lin):
        w = blob[lin.w_off: lin.w_off + lin.k * lin.n].astype(np.float64).reshape(lin.k, lin.n)
        return w, blob[lin.b_off: lin.b_off + lin.n].astype(np.float64)

    # head-major channel order: new channel h*32 + dd = reference channel 4*dd + h (models/superglue.py:110-113 views the
    # projection as [B, D/4, 4 heads, n]: the head index is the FAST part of the channel index)
    perm = np.array([4 * dd + h for h in range(4) for dd in range(D // 4)])
    words, biases = [], []

    def stages(w_kn):
        """[K, 128] -> list of K/64 stages (uint32 words of [hi|lo][128 rows][128 bytes])"""
        img = _sa_tc_images(w_kn)  # [K/64][hi|lo][N][64 fp16] flattened, SA_TC_WSCALE = 2^8
        return list(img.reshape(w_kn.shape[0] // 64, -1))

    mats = []
    for L in range(n_layers):
        wq, bq = mat(d.q[L])
        wk, bk = mat(d.k[L])
        wv, bv = mat(d.v[L])
        wm, bm = mat(d.merge[L])
        w0, b0 = mat(d.mlp0[L])
        w3, b3 = mat(d.mlp3[L])
        mats += [wq, wk, wv, wm, w0, w3]
        for w in (wq[:, perm], wk[:, perm], wv[:, perm], wm[perm, :]):
            words += stages(w)
        blocks = [stages(w0[:, j * 128:(j + 1) * 128]) for j in range(2)]  # [column block][K chunk 0..3]
        for kpair in ((0, 1), (2, 3)):
            for nb in range(2):
                words += [blocks[nb][kc] for kc in kpair]
        words += stages(w3)
        biases += [bq[perm], bk[perm], bv[perm], bm, b0, b3]
    wf, bf = mat(d.final_proj)
    mats.append(wf)
    words += stages(wf)
    biases.append(bf)
    if not all(fits_fp16_split(w, SG_TC_WSCALE) for w in mats):
        return -1, -1
    assert SG_TC_WSCALE == SA_TC_WSCALE and len(words) == n_layers * 20 + 2 and all(x.size == 8192 for x in words)
    return bb.add_raw_u32(np.concatenate(words)), bb.add(np.concatenate(biases))
