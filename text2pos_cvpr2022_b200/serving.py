"""Online coarse retrieval engine: tokens -> text embedding -> top-k cell indices against a resident DB.

This is the fast path ``bench.py`` measures: all device buffers are preallocated, the kernels of one step (device
tokeniser, tensor-core LSTM, finalize, top-k scan, select) are enqueued through the C ABI with no per-step allocation,
and the step can be captured once into a CUDA graph and replayed.  ``query(strings)`` is the end-to-end user call: the
raw bytes of the batch go into a pinned staging buffer (one native call that also enqueues the H2D copy), the step
runs, one D2H copy brings back scores, indices and the token counts (checked on the host).  Batches with non-ASCII
characters are tokenised by the Unicode-aware Python rules instead and skip the device tokeniser.

The engine owns ``depth`` independent *slots* (staging buffers, workspaces, a stream and CUDA graphs each).  With
``depth >= 2``, ``submit()`` / ``collect()`` keep several batches in flight: the host tokenises batch i+1 while the
GPU works on batch i, and the top-k of batch i overlaps the text encoder of batch i+1 on the idle SMs.
"""
import collections
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .modules import tokenize


class _Slot:
    """One in-flight batch: staging buffers both ways, the text embedding, workspaces, a stream and its graphs."""

    def __init__(self, eng: "OnlineRetrievalEngine", own_stream: bool):
        dev, B, T, k, D = eng.device, eng.B, eng.T, eng.k, eng.D
        # in: the staged text [offsets | bytes] (device tokeniser) -- or, for the host-tokenised fallback, tokens + lengths
        cap = eng.lib.t2p_stage_texts_capacity(B, B * eng.max_text_bytes)
        self.h_stage = torch.zeros(cap, dtype=torch.uint8).pin_memory()
        self.d_stage = torch.zeros(cap, dtype=torch.uint8, device=dev)
        self.tokens = torch.zeros(B, T, dtype=torch.int32, device=dev)
        self.h_tokens = torch.zeros(B, T, dtype=torch.int32).pin_memory()
        self.h_lengths = torch.ones(B, dtype=torch.int32).pin_memory()
        # out: one buffer [scores B*k f64 | idx B*k i64 | token counts B i32] -> one D2H copy
        n_len64 = (B + 1) // 2
        self.d_out = torch.zeros(2 * B * k + n_len64, dtype=torch.int64, device=dev)
        self.h_out = torch.zeros(2 * B * k + n_len64, dtype=torch.int64).pin_memory()
        self.out_scores = self.d_out[: B * k].view(torch.float64).view(B, k)
        self.out_idx = self.d_out[B * k: 2 * B * k].view(B, k)
        self.lengths = self.d_out[2 * B * k:].view(torch.int32)[:B]
        self.lengths.fill_(1)
        self.h_scores = self.h_out[: B * k].view(torch.float64).view(B, k)
        self.h_idx = self.h_out[B * k: 2 * B * k].view(B, k)
        self.h_counts = self.h_out[2 * B * k:].view(torch.int32)[:B].numpy()
        self.used_bytes = 0
        self.q = torch.empty(B, D, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            self.ws_lstm = torch.empty(max(256, eng.lib.t2p_lstm_encode_workspace(B, D)), dtype=torch.uint8, device=dev)
            self.stream = torch.cuda.Stream() if own_stream else None
            self.done = torch.cuda.Event()
        self.ws_topk = None
        self.graphs = {}


class OnlineRetrievalEngine:
    KERNELS_PER_STEP = 5  # tokenize, lstm_tc, lstm_finalize, retrieve_scan_tc, retrieve_select

    def __init__(self, model, db: torch.Tensor, k: int = 10, max_batch: int = 64, max_tokens: int = 64,
                 idx_base: int = 0, cell_ids: Optional[Sequence[str]] = None, depth: int = 1, max_text_bytes: int = 1024):
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
        self.vocab.to_device(self.device)
        self.max_text_bytes = int(max_text_bytes)  # average bytes per description the staging buffer is sized for
        self.depth = max(1, int(depth))
        # slot 0 runs on the caller's current stream (query / enqueue_*); further slots own a stream each
        self.slots = [_Slot(self, own_stream=(i > 0 or self.depth > 1)) for i in range(self.depth)]
        self._inflight = collections.deque()
        self._next = 0
        self.set_db(db)

    # slot 0 under the historical attribute names
    def __getattr__(self, name):
        if name in ("d_stage", "h_stage", "tokens", "lengths", "h_tokens", "h_lengths", "d_out", "h_out", "out_scores", "out_idx",
                    "h_scores", "h_idx", "q", "ws_lstm", "ws_topk"):
            return getattr(self.__dict__["slots"][0], name)
        if name == "_graphs":
            return self.__dict__["slots"][0].graphs
        raise AttributeError(name)

    def set_db(self, db: torch.Tensor):
        from .retrieval import db_row_norm2_max

        _lib.require_cuda(db, "cell database")
        self.db = db.float().contiguous()
        self.db_norm2_max = db_row_norm2_max(self.db)  # once per DB: certification bound of the tensor-core scan
        self.stats = torch.zeros(2, dtype=torch.int32, device=self.device)  # [certified, rescanned] query counters
        with torch.cuda.device(self.device):
            n = self.lib.t2p_retrieve_topk_workspace(self.B, self.db.shape[0], self.db.shape[1], self.k)
            for s in self.slots:
                s.ws_topk = torch.empty(max(256, n), dtype=torch.uint8, device=self.device)
                s.graphs = {}

    # ---- one step on the current stream ---------------------------------------------------------------------------
    def enqueue_encode(self, tokens: Optional[torch.Tensor] = None, lengths: Optional[torch.Tensor] = None, slot: int = 0):
        s = self.slots[slot]
        tokens = s.tokens if tokens is None else tokens
        lengths = s.lengths if lengths is None else lengths
        _lib.check(
            self.lib.t2p_lstm_encode(self.weights.handle, self.lstm_desc, tokens.data_ptr(), lengths.data_ptr(),
                                     self.B, tokens.shape[1], 1, s.q.data_ptr(), s.ws_lstm.data_ptr(), s.ws_lstm.numel(),
                                     _lib.stream_ptr(self.device)),
            "lstm_encode",
        )

    def enqueue_topk(self, db: Optional[torch.Tensor] = None, slot: int = 0):
        """``db``: an alternative resident copy with the SAME rows (hence the same norm bound) as ``self.db``."""
        s = self.slots[slot]
        db = self.db if db is None else db
        _lib.check(
            self.lib.t2p_retrieve_topk_ex(s.q.data_ptr(), db.data_ptr(), self.B, db.shape[0], db.shape[1], self.k,
                                          self.idx_base, self.db_norm2_max.data_ptr(), 0, s.out_scores.data_ptr(),
                                          s.out_idx.data_ptr(), self.stats.data_ptr(), s.ws_topk.data_ptr(),
                                          s.ws_topk.numel(), _lib.stream_ptr(self.device)),
            "retrieve_topk",
        )

    def enqueue_tokenize(self, slot: int = 0, d_stage: Optional[torch.Tensor] = None):
        """Device tokeniser: staged text (of ``slot``, or an alternative resident staging buffer) -> the slot's tokens / lengths."""
        s = self.slots[slot]
        self.vocab.tokenize_device(s.d_stage if d_stage is None else d_stage, self.B, s.tokens, s.lengths)

    def enqueue_step(self, db: Optional[torch.Tensor] = None, slot: int = 0, tokenize: bool = True):
        if tokenize:
            self.enqueue_tokenize(slot)
        self.enqueue_encode(slot=slot)
        self.enqueue_topk(db, slot=slot)

    def capture(self, key=0, db: Optional[torch.Tensor] = None, slot: int = 0):
        """Capture one step (staging buffers of ``slot`` -> top-k against ``db``) into a CUDA graph stored under ``key``."""
        with torch.cuda.device(self.device):
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self.enqueue_step(db, slot=slot)  # warm-up outside capture (cudaFuncSetAttribute etc.)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.enqueue_step(db, slot=slot)
            self.slots[slot].graphs[key] = g
        return g

    def capture_all(self, key=0, db: Optional[torch.Tensor] = None):
        for i in range(self.depth):
            self.capture(key, db, slot=i)

    def replay(self, key=0, slot: int = 0):
        self.slots[slot].graphs[key].replay()

    # ---- end-to-end user calls ------------------------------------------------------------------------------------
    def _stage(self, descriptions: Sequence[str], slot: int) -> bool:
        """Stage one batch for ``slot`` on the current stream.  ASCII batches: raw bytes -> pinned buffer -> H2D (one
        native call), tokenised on the device; returns True.  Otherwise: Python tokeniser (Unicode-aware lower()/split()),
        tokens + lengths copied to the device; returns False (the step then skips the device tokeniser)."""
        s = self.slots[slot]
        if len(descriptions) != self.B:
            raise ValueError(f"engine built for batches of {self.B} queries, got {len(descriptions)}")
        try:
            s.used_bytes, ascii_ = self.vocab.stage_texts(descriptions, s.h_stage, s.d_stage)
        except RuntimeError as e:
            raise ValueError(f"batch does not fit the staging buffer (raise max_text_bytes): {e}") from None
        if ascii_:
            return True
        tokens, lengths = tokenize(descriptions, self.known_words)
        if tokens.shape[1] > self.T:
            raise ValueError(f"engine built for <= {self.T} tokens per query, got {tokens.shape[1]}")
        if int(lengths.min()) < 1:
            raise ValueError("empty description (the reference's packed LSTM rejects length 0 too)")
        s.h_tokens.zero_()
        s.h_tokens[:, : tokens.shape[1]] = torch.from_numpy(tokens)
        s.h_lengths[:] = torch.from_numpy(lengths)
        s.tokens.copy_(s.h_tokens, non_blocking=True)
        s.lengths.copy_(s.h_lengths, non_blocking=True)
        s.used_bytes = s.h_tokens.numel() * 4 + s.h_lengths.numel() * 4
        return False

    def load_tokens(self, tokens: np.ndarray, lengths: np.ndarray):
        B, T = tokens.shape
        if B != self.B or T > self.T:
            raise ValueError(f"engine built for batch {self.B} x <= {self.T} tokens, got {B} x {T}")
        self.h_tokens.zero_()
        self.h_tokens[:, :T] = torch.from_numpy(tokens)
        self.h_lengths[:] = torch.from_numpy(lengths)
        self.tokens.copy_(self.h_tokens, non_blocking=True)
        self.lengths.copy_(self.h_lengths, non_blocking=True)

    def _enqueue_query(self, s: _Slot, slot: int, descriptions, graph_key):
        on_device = self._stage(descriptions, slot)
        if on_device and graph_key is not None and graph_key in s.graphs:
            s.graphs[graph_key].replay()
        else:
            self.enqueue_step(slot=slot, tokenize=on_device)
        s.h_out.copy_(s.d_out, non_blocking=True)

    def _check_counts(self, s: _Slot):
        c = s.h_counts
        lo, hi = int(c.min()), int(c.max())
        if lo < 1 or hi > self.T:
            if lo < 0:
                raise ValueError("a description is longer than the device tokeniser takes (8192 bytes)")
            if lo == 0:
                raise ValueError("empty description (the reference's packed LSTM rejects length 0 too)")
            raise ValueError(f"engine built for <= {self.T} tokens per query")

    def query(self, descriptions: List[str], graph_key=None):
        """strings -> (idx [B,k] int64 numpy, scores [B,k] float64 numpy); synchronous, slot 0.
        One pinned H2D copy, the five kernels (a captured CUDA graph if ``graph_key`` names one), one D2H copy."""
        if self._inflight:
            raise RuntimeError("query() while submitted batches are in flight: collect() them first")
        s = self.slots[0]
        if s.stream is not None:
            with torch.cuda.stream(s.stream):
                self._enqueue_query(s, 0, descriptions, graph_key)
            s.stream.synchronize()
        else:
            self._enqueue_query(s, 0, descriptions, graph_key)
            torch.cuda.current_stream(self.device).synchronize()
        self._check_counts(s)
        return s.h_idx.numpy(), s.h_scores.numpy()

    def submit(self, descriptions: List[str], graph_key=None) -> int:
        """Asynchronous ``query``: stages the batch into the next free slot, enqueues H2D + step + D2H on the slot's
        stream and returns the slot number.  At most ``depth`` batches may be in flight; results come back in order via
        ``collect``."""
        if self.depth < 2:
            raise RuntimeError("submit() needs an engine built with depth >= 2")
        if len(self._inflight) >= self.depth:
            raise RuntimeError(f"{self.depth} batches already in flight: collect() first")
        slot = self._next
        s = self.slots[slot]
        with torch.cuda.stream(s.stream):
            self._enqueue_query(s, slot, descriptions, graph_key)
            s.done.record()
        self._next = (self._next + 1) % self.depth
        self._inflight.append(slot)
        return slot

    def collect(self):
        """Result of the oldest submitted batch: (idx, scores) numpy views of its pinned buffers, valid until the slot
        is submitted again."""
        slot = self._inflight.popleft()
        s = self.slots[slot]
        s.done.synchronize()
        self._check_counts(s)
        return s.h_idx.numpy(), s.h_scores.numpy()

    def h2d_bytes(self) -> int:
        """Bytes of the last staged batch of slot 0 (raw text + offsets; they vary with the text)."""
        return int(self.slots[0].used_bytes)

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
        self.dist.all_gather_into_tensor(self.gathered.view(self.world * 2, e.B, e.k), e.d_out[: 2 * e.B * e.k].view(2, e.B, e.k), group=self.group)
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
        on_device = e._stage(descriptions, 0)
        if on_device and graph_key is not None and graph_key in e._graphs:
            e._graphs[graph_key].replay()
        else:
            e.enqueue_step(tokenize=on_device)
        self.enqueue_exchange()
        self.h_final.copy_(self.d_final, non_blocking=True)
        torch.cuda.current_stream(e.device).synchronize()
        B, k = e.B, e.k
        return self.h_final[B * k:].view(B, k).numpy(), self.h_final[: B * k].view(torch.float64).view(B, k).numpy()
