"""All-pairs scores + top-k over a resident cell database (reference: ``training/coarse.py:134-148``).

``CellDatabase`` keeps the ``[N, D]`` float32 cell embeddings (and the cell-id strings) resident in HBM;
``topk`` runs ``t2p_retrieve_topk``.  ``ShardedCellDatabase`` partitions the rows over the ranks of a
``torch.distributed`` group (rank r owns rows ``[r*ceil(N/R), ...)``), runs the local top-k, exchanges the
per-shard lists with ONE all-gather (NCCL over NVLink on GPUs) and merges them with ``t2p_topk_merge``
(``topk``: queries replicated; ``topk_dp``: every rank brings its own query batch, one more all-gather).
Ordering everywhere: (float64 score descending, index ascending).
"""
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def db_row_norm2_max(db: torch.Tensor) -> torch.Tensor:
    """Device scalar (float32 [1]) = max squared row norm of ``db``: the certification bound of the tensor-core scan.
    Computed once per database (``t2p_db_row_norm2_max``)."""
    lib = _lib.load()
    _lib.require_cuda(db, "cell database")
    db = db.float().contiguous()
    out = torch.zeros(1, dtype=torch.float32, device=db.device)
    with torch.cuda.device(db.device):
        _lib.check(lib.t2p_db_row_norm2_max(_lib.ptr(db), db.shape[0], db.shape[1], _lib.ptr(out), _lib.stream_ptr(db.device)),
                   "db_row_norm2_max")
    return out


def retrieve_topk(q: torch.Tensor, db: torch.Tensor, k: int, idx_base: int = 0, workspace: Optional[_lib.Workspace] = None,
                  norm2_max: Optional[torch.Tensor] = None, flags: int = 0, stats: Optional[torch.Tensor] = None):
    """q [B,D], db [N,D] float32 CUDA -> (idx [B,k] int64, scores [B,k] float64).

    ``norm2_max``: optional result of :func:`db_row_norm2_max` (saves one pass over the DB per call on the tensor-core
    path); ``flags``: ``_lib.RETRIEVE_FORCE_*``; ``stats``: optional int32 [2] device counters (certified, rescanned)."""
    lib = _lib.load()
    _lib.require_cuda(q, "queries")
    _lib.require_cuda(db, "cell database")
    q = q.float().contiguous()
    db = db.float().contiguous()
    B, D = q.shape
    N = db.shape[0]
    if db.shape[1] != D:
        raise ValueError(f"retrieve_topk: query dim {D} != database dim {db.shape[1]}")
    dev = q.device
    out_s = torch.empty(B, k, dtype=torch.float64, device=dev)
    out_i = torch.empty(B, k, dtype=torch.int64, device=dev)
    if B == 0:
        return out_i, out_s
    ws_owner = workspace if workspace is not None else _lib.Workspace()
    with torch.cuda.device(dev):
        ws = ws_owner.get(lib.t2p_retrieve_topk_workspace(B, N, D, k), dev)
        _lib.check(
            lib.t2p_retrieve_topk_ex(_lib.ptr(q), _lib.ptr(db), B, N, D, k, int(idx_base), _lib.ptr(norm2_max), int(flags),
                                     _lib.ptr(out_s), _lib.ptr(out_i), _lib.ptr(stats), _lib.ptr(ws), ws.numel(),
                                     _lib.stream_ptr(dev)),
            "retrieve_topk",
        )
    return out_i, out_s


def topk_merge(scores: torch.Tensor, idx: torch.Tensor, k_out: int):
    """scores [R,B,k_in] float64, idx [R,B,k_in] int64 (global indices, -1 = empty) -> ([B,k_out] idx, scores)."""
    lib = _lib.load()
    _lib.require_cuda(scores, "shard scores")
    scores = scores.contiguous()
    idx = idx.contiguous()
    R, B, k_in = scores.shape
    dev = scores.device
    out_s = torch.empty(B, k_out, dtype=torch.float64, device=dev)
    out_i = torch.empty(B, k_out, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(
            lib.t2p_topk_merge(_lib.ptr(scores), _lib.ptr(idx), R, B, k_in, k_out, _lib.ptr(out_s), _lib.ptr(out_i),
                               _lib.stream_ptr(dev)),
            "topk_merge",
        )
    return out_i, out_s


class CellDatabase:
    """Resident embeddings + ids of the cells of one shard (or of the whole DB)."""

    def __init__(self, embeddings: torch.Tensor, cell_ids: Optional[Sequence[str]] = None, idx_base: int = 0):
        self.embeddings = embeddings.float().contiguous()
        self.cell_ids = None if cell_ids is None else np.asarray(cell_ids)
        self.idx_base = int(idx_base)
        self._ws = _lib.Workspace()
        self.norm2_max = db_row_norm2_max(self.embeddings) if len(self) else None

    def __len__(self):
        return self.embeddings.shape[0]

    def topk(self, queries: torch.Tensor, k: int):
        return retrieve_topk(queries, self.embeddings, k, self.idx_base, self._ws, self.norm2_max)

    def topk_ids(self, queries: torch.Tensor, k: int) -> np.ndarray:
        """-> [B,k] array of cell-id strings, what ``eval_epoch`` stores in ``top_retrievals``."""
        assert self.cell_ids is not None
        idx, _ = self.topk(queries, min(int(k), len(self)))  # never index the ids with the -1 padding of a short DB
        return self.cell_ids[(idx - self.idx_base).cpu().numpy()]


def shard_bounds(n: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous row blocks: rank r owns [r*ceil(n/R), min(n, (r+1)*ceil(n/R)))."""
    per = (n + world_size - 1) // world_size
    return [(min(n, r * per), min(n, (r + 1) * per)) for r in range(world_size)]


class ShardedCellDatabase:
    """One shard per rank; ``topk`` = local top-k -> all-gather -> merge.

    ``local_topk`` / ``merge`` are injectable so that the exchange logic can be exercised on CPU ``gloo`` groups in
    the tests (with the oracle as the stand-in); the defaults are the CUDA kernels and there is no CPU fallback.
    """

    def __init__(self, local_embeddings: torch.Tensor, n_total: int, group=None,
                 local_topk: Optional[Callable] = None, merge: Optional[Callable] = None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.n_total = int(n_total)
        self.lo, self.hi = shard_bounds(self.n_total, self.world)[self.rank]
        if local_embeddings.shape[0] != self.hi - self.lo:
            raise ValueError(f"rank {self.rank} owns rows [{self.lo},{self.hi}) but got {local_embeddings.shape[0]} rows")
        self.local = local_embeddings
        self._local_topk = local_topk
        self._merge = merge
        self._ws = None if local_topk is not None else _lib.Workspace()
        self._norm2_max = db_row_norm2_max(local_embeddings) if (local_topk is None and self.hi > self.lo) else None

    def topk(self, queries: torch.Tensor, k: int):
        """queries [B,D] replicated on every rank -> (idx [B,k] global int64, scores [B,k] float64) on every rank."""
        B = queries.shape[0]
        dev = queries.device
        if self.hi > self.lo:
            if self._local_topk is not None:
                li, ls = self._local_topk(queries, self.local, k, self.lo)
            else:
                li, ls = retrieve_topk(queries, self.local, k, self.lo, self._ws, self._norm2_max)
        else:  # empty shard
            li = torch.full((B, k), -1, dtype=torch.int64, device=dev)
            ls = torch.full((B, k), float("-inf"), dtype=torch.float64, device=dev)
        # one exchange: (score, index) pairs packed into a single int64 buffer -> ONE all-gather
        packed = torch.stack([ls.view(torch.int64), li], dim=0).contiguous()  # [2,B,k]
        flat = torch.empty((self.world * 2, B, k), dtype=torch.int64, device=dev)  # concatenation along dim 0
        self.dist.all_gather_into_tensor(flat, packed, group=self.group)
        gathered = flat.view(self.world, 2, B, k)
        gs = gathered[:, 0].contiguous().view(torch.float64)  # [R,B,k]
        gi = gathered[:, 1].contiguous()
        if self._merge is not None:
            return self._merge(gs, gi, k)
        return topk_merge(gs, gi, k)

    def topk_dp(self, local_queries: torch.Tensor, k: int):
        """Data-parallel queries over the sharded DB: every rank passes ITS OWN [B,D] batch (same B on every rank) and
        gets the global top-k of its own queries.  all-gather of the query embeddings (B*D*4 bytes per rank) -> local
        top-k of all R*B queries against the shard -> all-gather of the per-shard lists -> merge of the own rows."""
        B, D = local_queries.shape
        R, dev = self.world, local_queries.device
        q_all = torch.empty((R * B, D), dtype=local_queries.dtype, device=dev)
        self.dist.all_gather_into_tensor(q_all, local_queries.contiguous(), group=self.group)
        if self.hi > self.lo:
            if self._local_topk is not None:
                li, ls = self._local_topk(q_all, self.local, k, self.lo)
            else:
                li, ls = retrieve_topk(q_all, self.local, k, self.lo, self._ws, self._norm2_max)
        else:  # empty shard
            li = torch.full((R * B, k), -1, dtype=torch.int64, device=dev)
            ls = torch.full((R * B, k), float("-inf"), dtype=torch.float64, device=dev)
        packed = torch.stack([ls.view(torch.int64), li], dim=0).contiguous()  # [2, R*B, k]
        flat = torch.empty((R * 2, R * B, k), dtype=torch.int64, device=dev)
        self.dist.all_gather_into_tensor(flat, packed, group=self.group)
        mine = flat.view(R, 2, R, B, k)[:, :, self.rank]  # [shard, {score,idx}, B, k]: the rows of this rank's queries
        gs = mine[:, 0].contiguous().view(torch.float64)
        gi = mine[:, 1].contiguous()
        if self._merge is not None:
            return self._merge(gs, gi, k)
        return topk_merge(gs, gi, k)
