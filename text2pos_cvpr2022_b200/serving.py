"""Online coarse retrieval engine: tokens -> text embedding -> top-k cell indices against a resident DB.

This is the fast path ``bench.py`` measures: all device buffers are preallocated, the four kernels of one step
(cluster LSTM, finalize, partial top-k, merge) are enqueued through the C ABI with no per-step allocation, and the
step can be captured once into a CUDA graph and replayed.  ``query(strings)`` is the end-to-end user call: host
tokenisation, pinned staging, H2D, step, D2H.
"""
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .modules import tokenize


class OnlineRetrievalEngine:
    KERNELS_PER_STEP = 4  # lstm_reg, lstm_finalize, retrieve_scan_tc, retrieve_select

    def __init__(self, model, db: torch.Tensor, k: int = 10, max_batch: int = 64, max_tokens: int = 64,
                 idx_base: int = 0, cell_ids: Optional[Sequence[str]] = None):
        self.lib = _lib.load()
        self.model = model
        self.weights, desc = model.t2p_packed()
        self.lstm_desc = desc["lstm"] if isinstance(desc, dict) else desc
        self.device = model.t2p_device()
        self.k, self.B, self.T = int(k), int(max_batch), int(max_tokens)
        self.D = self.lstm_desc.hidden
        self.idx_base = int(idx_base)
        self.cell_ids = None if cell_ids is None else np.asarray(cell_ids)
        self.known_words = model.language_encoder.known_words if hasattr(model, "language_encoder") else model.known_words
        self.vocab = _lib.Vocab(self.known_words)
        self.set_db(db)
        dev = self.device
        B, T, k = self.B, self.T, self.k
        # one staging buffer each way: [tokens B*T | lengths B] int32 in, [scores B*k f64 | idx B*k i64] out
        self.d_in = torch.zeros(B * T + B, dtype=torch.int32, device=dev)
        self.d_in[B * T:] = 1
        self.h_in = torch.zeros(B * T + B, dtype=torch.int32).pin_memory()
        self.h_in[B * T:] = 1
        self.tokens = self.d_in[: B * T].view(B, T)
        self.lengths = self.d_in[B * T:]
        self.h_tokens = self.h_in[: B * T].view(B, T)
        self.h_lengths = self.h_in[B * T:]
        self.d_out = torch.empty(2 * B * k, dtype=torch.int64, device=dev)
        self.h_out = torch.empty(2 * B * k, dtype=torch.int64).pin_memory()
        self.out_scores = self.d_out[: B * k].view(torch.float64).view(B, k)
        self.out_idx = self.d_out[B * k:].view(B, k)
        self.h_scores = self.h_out[: B * k].view(torch.float64).view(B, k)
        self.h_idx = self.h_out[B * k:].view(B, k)
        self.q = torch.empty(B, self.D, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            self.ws_lstm = torch.empty(max(256, self.lib.t2p_lstm_encode_workspace(B, self.D)), dtype=torch.uint8, device=dev)
        self._graphs = {}

    def set_db(self, db: torch.Tensor):
        from .retrieval import db_row_norm2_max

        _lib.require_cuda(db, "cell database")
        self.db = db.float().contiguous()
        self.db_norm2_max = db_row_norm2_max(self.db)  # once per DB: certification bound of the tensor-core scan
        self.stats = torch.zeros(2, dtype=torch.int32, device=self.device)  # [certified, rescanned] query counters
        with torch.cuda.device(self.device):
            n = self.lib.t2p_retrieve_topk_workspace(self.B, self.db.shape[0], self.db.shape[1], self.k)
            self.ws_topk = torch.empty(max(256, n), dtype=torch.uint8, device=self.device)
        self._graphs = {}

    # ---- one step on the current stream ---------------------------------------------------------------------------
    def enqueue_encode(self, tokens: Optional[torch.Tensor] = None, lengths: Optional[torch.Tensor] = None):
        tokens = self.tokens if tokens is None else tokens
        lengths = self.lengths if lengths is None else lengths
        _lib.check(
            self.lib.t2p_lstm_encode(self.weights.handle, self.lstm_desc, tokens.data_ptr(), lengths.data_ptr(),
                                     self.B, tokens.shape[1], 1, self.q.data_ptr(), self.ws_lstm.data_ptr(), self.ws_lstm.numel(),
                                     _lib.stream_ptr(self.device)),
            "lstm_encode",
        )

    def enqueue_topk(self, db: Optional[torch.Tensor] = None):
        """``db``: an alternative resident copy with the SAME rows (hence the same norm bound) as ``self.db``."""
        db = self.db if db is None else db
        _lib.check(
            self.lib.t2p_retrieve_topk_ex(self.q.data_ptr(), db.data_ptr(), self.B, db.shape[0], db.shape[1], self.k,
                                          self.idx_base, self.db_norm2_max.data_ptr(), 0, self.out_scores.data_ptr(),
                                          self.out_idx.data_ptr(), self.stats.data_ptr(), self.ws_topk.data_ptr(),
                                          self.ws_topk.numel(), _lib.stream_ptr(self.device)),
            "retrieve_topk",
        )

    def enqueue_step(self, db: Optional[torch.Tensor] = None):
        self.enqueue_encode()
        self.enqueue_topk(db)

    def capture(self, key=0, db: Optional[torch.Tensor] = None):
        """Capture one step (staging buffers -> top-k against ``db``) into a CUDA graph stored under ``key``."""
        with torch.cuda.device(self.device):
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self.enqueue_step(db)  # warm-up outside capture (cudaFuncSetAttribute etc.)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.enqueue_step(db)
            self._graphs[key] = g
        return g

    def replay(self, key=0):
        self._graphs[key].replay()

    # ---- end-to-end user call -------------------------------------------------------------------------------------
    def stage_queries(self, descriptions: Sequence[str]):
        """Host tokenisation straight into the pinned staging buffer (native tokeniser; Unicode strings take the
        Python one, whose lower()/split() are Unicode-aware)."""
        if len(descriptions) != self.B:
            raise ValueError(f"engine built for batches of {self.B} queries, got {len(descriptions)}")
        if all(d.isascii() for d in descriptions):
            self.vocab.tokenize_into(descriptions, self.h_tokens, self.h_lengths)
        else:
            tokens, lengths = tokenize(descriptions, self.known_words)
            if tokens.shape[1] > self.T:
                raise ValueError(f"engine built for <= {self.T} tokens per query, got {tokens.shape[1]}")
            self.h_tokens.zero_()
            self.h_tokens[:, : tokens.shape[1]] = torch.from_numpy(tokens)
            self.h_lengths[:] = torch.from_numpy(lengths)
        if int(self.h_lengths.min()) < 1:
            raise ValueError("empty description (the reference's packed LSTM rejects length 0 too)")

    def load_tokens(self, tokens: np.ndarray, lengths: np.ndarray):
        B, T = tokens.shape
        if B != self.B or T > self.T:
            raise ValueError(f"engine built for batch {self.B} x <= {self.T} tokens, got {B} x {T}")
        self.h_tokens.zero_()
        self.h_tokens[:, :T] = torch.from_numpy(tokens)
        self.h_lengths[:] = torch.from_numpy(lengths)
        self.d_in.copy_(self.h_in, non_blocking=True)

    def query(self, descriptions: List[str], graph_key=None):
        """strings -> (idx [B,k] int64 numpy, scores [B,k] float64 numpy); synchronous.
        One pinned H2D copy, the four kernels (a captured CUDA graph if ``graph_key`` names one), one D2H copy."""
        self.stage_queries(descriptions)
        self.d_in.copy_(self.h_in, non_blocking=True)
        if graph_key is not None and graph_key in self._graphs:
            self._graphs[graph_key].replay()
        else:
            self.enqueue_step()
        self.h_out.copy_(self.d_out, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self.h_idx.numpy(), self.h_scores.numpy()

    def h2d_bytes(self) -> int:
        return self.h_in.numel() * 4

    def d2h_bytes(self) -> int:
        return self.h_out.numel() * 8


class ShardedOnlineRetrievalEngine:
    """One ``OnlineRetrievalEngine`` per rank over its row shard of the DB (``idx_base`` = first row); queries are
    replicated.  ``step``/``query`` = local step -> ONE all-gather of the packed [2, B, k] int64 result (score bits,
    global index; 10 KB per rank at B=64, k=10) -> ``t2p_topk_merge`` on every rank."""

    def __init__(self, engine: OnlineRetrievalEngine, group=None):
        import torch.distributed as dist

        self.dist, self.group, self.eng = dist, group, engine
        self.world = dist.get_world_size(group)
        B, k, dev = engine.B, engine.k, engine.device
        self.gathered = torch.empty((self.world, 2, B, k), dtype=torch.int64, device=dev)
        self.d_final = torch.empty(2 * B * k, dtype=torch.int64, device=dev)
        self.h_final = torch.empty(2 * B * k, dtype=torch.int64).pin_memory()
        self.final_scores = self.d_final[: B * k].view(torch.float64).view(B, k)
        self.final_idx = self.d_final[B * k:].view(B, k)

    def enqueue_exchange(self):
        """all-gather + merge of the local result that ``engine.enqueue_step`` left in ``engine.d_out``."""
        e = self.eng
        self.dist.all_gather_into_tensor(self.gathered.view(self.world * 2, e.B, e.k), e.d_out.view(2, e.B, e.k), group=self.group)
        # gathered[r, 0] / [r, 1] are contiguous [B, k] blocks: the merge kernel takes the two [R, B, k] arrays as strided
        # views only if contiguous, so split once (2 x R*B*k*8 bytes, device-to-device)
        gs = self.gathered[:, 0].contiguous().view(torch.float64)
        gi = self.gathered[:, 1].contiguous()
        _lib.check(
            e.lib.t2p_topk_merge(gs.data_ptr(), gi.data_ptr(), self.world, e.B, e.k, e.k, self.final_scores.data_ptr(),
                                 self.final_idx.data_ptr(), _lib.stream_ptr(e.device)),
            "topk_merge",
        )

    def query(self, descriptions: List[str], graph_key=None):
        e = self.eng
        e.stage_queries(descriptions)
        e.d_in.copy_(e.h_in, non_blocking=True)
        if graph_key is not None and graph_key in e._graphs:
            e._graphs[graph_key].replay()
        else:
            e.enqueue_step()
        self.enqueue_exchange()
        self.h_final.copy_(self.d_final, non_blocking=True)
        torch.cuda.current_stream(e.device).synchronize()
        B, k = e.B, e.k
        return self.h_final[B * k:].view(B, k).numpy(), self.h_final[: B * k].view(torch.float64).view(B, k).numpy()
