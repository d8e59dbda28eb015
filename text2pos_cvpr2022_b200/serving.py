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
    KERNELS_PER_STEP = 4  # lstm_cluster, lstm_finalize, retrieve_partial, retrieve_merge

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
        self.set_db(db)
        dev = self.device
        self.tokens = torch.zeros(self.B, self.T, dtype=torch.int32, device=dev)
        self.lengths = torch.ones(self.B, dtype=torch.int32, device=dev)
        self.q = torch.empty(self.B, self.D, dtype=torch.float32, device=dev)
        self.out_idx = torch.empty(self.B, self.k, dtype=torch.int64, device=dev)
        self.out_scores = torch.empty(self.B, self.k, dtype=torch.float64, device=dev)
        self.h_tokens = torch.zeros(self.B, self.T, dtype=torch.int32).pin_memory()
        self.h_lengths = torch.ones(self.B, dtype=torch.int32).pin_memory()
        self.h_idx = torch.empty(self.B, self.k, dtype=torch.int64).pin_memory()
        self.h_scores = torch.empty(self.B, self.k, dtype=torch.float64).pin_memory()
        with torch.cuda.device(dev):
            self.ws_lstm = torch.empty(max(256, self.lib.t2p_lstm_encode_workspace(self.B, self.D)), dtype=torch.uint8, device=dev)
        self._graphs = {}

    def set_db(self, db: torch.Tensor):
        _lib.require_cuda(db, "cell database")
        self.db = db.float().contiguous()
        from .retrieval import db_row_norm2_max

        self.db_norm2_max = db_row_norm2_max(self.db)  # once per DB: certification bound of the tensor-core scan
        self.stats = torch.zeros(2, dtype=torch.int32, device=self.device)  # [certified, rescanned] query counters
        with torch.cuda.device(self.device):
            n = self.lib.t2p_retrieve_topk_workspace(self.B, self.db.shape[0], self.db.shape[1], self.k)
            self.ws_topk = torch.empty(max(256, n), dtype=torch.uint8, device=self.device)

    # ---- one step on the current stream, inputs already in self.tokens / self.lengths ------------------------
    def enqueue_encode(self):
        _lib.check(
            self.lib.t2p_lstm_encode(self.weights.handle, self.lstm_desc, self.tokens.data_ptr(), self.lengths.data_ptr(),
                                     self.B, self.T, 1, self.q.data_ptr(), self.ws_lstm.data_ptr(), self.ws_lstm.numel(),
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
        """Capture one step (against ``db``) into a CUDA graph stored under ``key``."""
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

    # ---- end-to-end user call ---------------------------------------------------------------------------------
    def load_tokens(self, tokens: np.ndarray, lengths: np.ndarray):
        B, T = tokens.shape
        if B != self.B or T > self.T:
            raise ValueError(f"engine built for batch {self.B} x <= {self.T} tokens, got {B} x {T}")
        self.h_tokens[:, :T] = torch.from_numpy(tokens)
        self.h_lengths[:] = torch.from_numpy(lengths)
        self.tokens.copy_(self.h_tokens, non_blocking=True)
        self.lengths.copy_(self.h_lengths, non_blocking=True)

    def query(self, descriptions: List[str], use_graph: bool = True):
        """strings -> (idx [B,k] int64 numpy, scores [B,k] float64 numpy); synchronous."""
        tokens, lengths = tokenize(descriptions, self.model.language_encoder.known_words if hasattr(self.model, "language_encoder") else self.model.known_words)
        self.load_tokens(tokens, lengths)
        if use_graph and 0 in self._graphs:
            self.replay(0)
        else:
            self.enqueue_step()
        self.h_idx.copy_(self.out_idx, non_blocking=True)
        self.h_scores.copy_(self.out_scores, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self.h_idx.numpy(), self.h_scores.numpy()

    def h2d_bytes(self) -> int:
        return self.h_tokens.numel() * 4 + self.h_lengths.numel() * 4

    def d2h_bytes(self) -> int:
        return self.h_idx.numel() * 8 + self.h_scores.numel() * 8
