"""Online coarse retrieval engine: tokens -> text embedding -> top-k cell indices against a resident DB.

This is the fast path ``bench.py`` measures: all device buffers are preallocated, the kernels of one step (device
tokeniser, tensor-core LSTM, finalize, top-k scan, select) are enqueued through the C ABI with no per-step allocation,
and the step can be captured once into a CUDA graph and replayed.  ``query(strings)`` is the end-to-end user call: the
raw bytes of the batch go into a pinned staging buffer (one native call that also enqueues the H2D copy), the step
runs, one D2H copy brings back scores, indices and the token counts (checked on the host).  Batches with non-ASCII
characters are tokenised by the Unicode-aware Python rules instead and skip the device tokeniser.

The engine owns ``depth`` independent *slots* (staging buffers, workspaces, a stream and CUDA graphs each).  With
``depth >= 2``, ``submit()`` / ``collect()`` keep several batches in flight: the host stages batch i+1 while the
GPU works on batch i, and the top-k of batch i overlaps the text encoder of batch i+1 on the idle SMs.
"""
import collections
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .modules import tokenize


def _native_submit(lib, descriptions, B, st, key, graph, d_out, h_out, event, stream) -> Optional[bool]:
    """One foreign call: stage + H2D + graph launch + D2H + event record on ``stream`` (``t2p_serving_submit``).
    Returns True if the batch was launched, False if it has non-ASCII bytes (nothing enqueued).  The raw handles of
    (slot, graph) are looked up once and cached: the per-batch host work is one join/encode and one ctypes call."""
    if len(descriptions) != B:
        raise ValueError(f"engine built for batches of {B} queries, got {len(descriptions)}")
    raw = st.raw.get(key)
    if raw is None:
        raw = st.raw[key] = (graph.raw_cuda_graph_exec(), d_out.data_ptr(), h_out.data_ptr(), h_out.numel() * 8,
                             event.cuda_event, stream.cuda_stream)
    blob = ("\0".join(descriptions) + "\0").encode("utf-8")
    used, ascii_ = st.c_used, st.c_ascii
    rc = lib.t2p_serving_submit(blob, len(blob), B, st.h_stage_ptr, st.stage_cap, st.d_stage_ptr, raw[0], raw[1], raw[2], raw[3],
                                raw[4], raw[5], st.c_used_ref, st.c_ascii_ref)
    if rc == -3:
        raise ValueError("batch does not fit the staging buffer (raise max_text_bytes)")
    if rc != 0:
        _lib.check(rc, "serving_submit")
    st.used_bytes = used.value
    return ascii_.value != 0


def _replay_plan(graph_slots, stream_slots, device, keys, first_slot=0):
    """ctypes arrays for ``t2p_serving_replay_many``: step i = graph ``keys[i]`` of slot ``(first_slot + i) % depth`` on that
    slot's stream (built once, replayed many times)."""
    n, depth = len(keys), len(graph_slots)
    execs, streams = (_lib.C.c_void_p * n)(), (_lib.C.c_void_p * n)()
    for i, key in enumerate(keys):
        sl = (first_slot + i) % depth
        execs[i] = graph_slots[sl].graphs[key].raw_cuda_graph_exec()
        st = stream_slots[sl].stream
        streams[i] = st.cuda_stream if st is not None else _lib.stream_ptr(device)
    return execs, streams, n


class _Slot:
    """One in-flight batch: staging buffers both ways, the text embedding, workspaces, a stream and its graphs."""

    def __init__(self, eng: "OnlineRetrievalEngine", own_stream: bool):
        dev, B, T, k, D = eng.device, eng.B, eng.T, eng.k, eng.D
        # in: the staged text [offsets | bytes] (device tokeniser) -- or, for the host-tokenised fallback, tokens + lengths
        cap = eng.lib.t2p_stage_texts_capacity(B, B * eng.max_text_bytes)
        self.h_stage = torch.zeros(cap, dtype=torch.uint8).pin_memory()
        self.d_stage = torch.zeros(cap, dtype=torch.uint8, device=dev)
        self.h_stage_ptr, self.d_stage_ptr, self.stage_cap = self.h_stage.data_ptr(), self.d_stage.data_ptr(), int(cap)
        self.raw = {}  # graph key -> raw handles for t2p_serving_submit
        self.c_used, self.c_ascii = _lib._SZ(0), _lib._I(0)
        self.c_used_ref, self.c_ascii_ref = _lib.C.byref(self.c_used), _lib.C.byref(self.c_ascii)
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
        self.np_idx, self.np_scores = self.h_idx.numpy(), self.h_scores.numpy()  # views of the pinned result buffer
        self.used_bytes = 0
        self.q = torch.empty(B, D, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            self.ws_lstm = torch.empty(max(256, eng.lib.t2p_lstm_encode_workspace(B, D)), dtype=torch.uint8, device=dev)
            self.stream = torch.cuda.Stream() if own_stream else None
            self.done = torch.cuda.Event()
            self.done.record()  # creates the underlying cudaEvent_t (its raw handle goes to t2p_serving_submit)
        self.ws_topk = None
        self.graphs = {}


class OnlineRetrievalEngine:
    KERNELS_PER_STEP = 6  # tokenize, lstm_tc, lstm_finalize, retrieve_scan_tc, retrieve_select_warp, retrieve_select (flagged only)

    def __init__(self, model, db: torch.Tensor, k: int = 10, max_batch: int = 64, max_tokens: int = 64,
                 idx_base: int = 0, cell_ids: Optional[Sequence[str]] = None, depth: int = 1, max_text_bytes: int = 1024,
                 lstm_clusters: Optional[int] = None, scan_ctas: Optional[int] = None):
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
        # LSTM clusters per direction: default 7 (lowest latency) for a synchronous engine; a pipelined one packs the batch
        # into 2 clusters (two ping-pong groups each) or, with >= 8 batches in flight, into 1 cluster of four groups:
        # more latency per batch, a fraction of the SM-time, so more batches share the chip
        if lstm_clusters is None:
            lstm_clusters = 0 if self.depth == 1 else (2 if self.depth < 8 else 1)
        if lstm_clusters:
            import copy

            self.lstm_desc = copy.copy(self.lstm_desc)
            self.lstm_desc.max_groups = int(lstm_clusters)
        # top-k scan CTAs: one per SM for a synchronous engine (lowest latency); a pipelined one caps them -- every CTA then
        # streams several DB tiles against its resident query tile and the select reads fewer key lists: a fraction of the
        # SM-time per batch (T2P_RETRIEVE_MAX_CTAS)
        if scan_ctas is None:
            scan_ctas = 0 if self.depth == 1 else 24
        self.scan_ctas = max(0, min(255, int(scan_ctas)))
        self.topk_flags = self.scan_ctas << 8
        # slot 0 runs on the caller's current stream (query / enqueue_*); further slots own a stream each
        self.slots = [_Slot(self, own_stream=(i > 0 or self.depth > 1)) for i in range(self.depth)]
        self._inflight = collections.deque()
        self._next = 0
        self.set_db(db)

    # ---- slot 0 (the synchronous `query` path) by name ---------------------------------------------------------------
    @property
    def d_stage(self):
        return self.slots[0].d_stage

    @property
    def h_stage(self):
        return self.slots[0].h_stage

    @property
    def tokens(self):
        return self.slots[0].tokens

    @property
    def lengths(self):
        return self.slots[0].lengths

    @property
    def h_tokens(self):
        return self.slots[0].h_tokens

    @property
    def h_lengths(self):
        return self.slots[0].h_lengths

    @property
    def d_out(self):
        return self.slots[0].d_out

    @property
    def h_out(self):
        return self.slots[0].h_out

    @property
    def out_scores(self):
        return self.slots[0].out_scores

    @property
    def out_idx(self):
        return self.slots[0].out_idx

    @property
    def h_scores(self):
        return self.slots[0].h_scores

    @property
    def h_idx(self):
        return self.slots[0].h_idx

    @property
    def q(self):
        return self.slots[0].q

    @property
    def ws_lstm(self):
        return self.slots[0].ws_lstm

    @property
    def ws_topk(self):
        return self.slots[0].ws_topk

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
                s.raw = {}

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
                                          self.idx_base, self.db_norm2_max.data_ptr(), self.topk_flags, s.out_scores.data_ptr(),
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
            self.slots[slot].raw.pop(key, None)
        return g

    def capture_all(self, key=0, db: Optional[torch.Tensor] = None):
        for i in range(self.depth):
            self.capture(key, db, slot=i)

    def replay(self, key=0, slot: int = 0):
        self.slots[slot].graphs[key].replay()

    def replay_plan(self, keys: Sequence, first_slot: int = 0):
        """Plan for ``replay_many``: step i replays the captured graph ``keys[i]`` on slot ``(first_slot + i) % depth``."""
        return _replay_plan(self.slots, self.slots, self.device, keys, first_slot)

    def replay_many(self, plan, fork_join: bool = True):
        """Launch every step of the plan back to back from native code (device-resident inputs; no interpreter between the
        launches).  ``fork_join``: the slots' streams first wait for the current stream and the current stream waits for
        them at the end (events recorded around the call bracket the region); otherwise synchronise the slots yourself."""
        _lib.check(self.lib.t2p_serving_replay_many(*plan, _lib.stream_ptr(self.device), int(fork_join)), "serving_replay_many")

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
        g = s.graphs.get(graph_key) if graph_key is not None else None
        if g is None or not _native_submit(self.lib, descriptions, self.B, s, graph_key, g, s.d_out, s.h_out, s.done, s.stream):
            with torch.cuda.stream(s.stream):  # no graph for this key, or a non-ASCII batch: the general path
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
        return s.np_idx, s.np_scores

    def h2d_bytes(self) -> int:
        """Bytes of the last staged batch of slot 0 (raw text + offsets; they vary with the text)."""
        return int(self.slots[0].used_bytes)

    def d2h_bytes(self) -> int:
        return self.h_out.numel() * 8


class _ShardSlot:
    def __init__(self, sh: "ShardedOnlineRetrievalEngine"):
        e, R = sh.eng, sh.world
        B, k, D, dev = e.B, e.k, e.D, e.device
        # symmetric region (mapped by the peers in p2p mode): [q_all R*B*D f32 | mine 2*R*B*k i64 | flags 2*R u64]
        al = lambda x: (x + 255) // 256 * 256
        self.q_all_off, q_bytes = 0, R * B * D * 4
        self.mine_off, m_bytes = al(q_bytes), 2 * R * B * k * 8
        self.flags_off = al(self.mine_off + m_bytes)
        sym_bytes = self.flags_off + al(2 * R * 8)
        if sh.exchange == "p2p":  # a cudaMalloc block with a CUDA IPC handle (kernels of the peers store into it over NVLink)
            self.sym_buf = _lib.SymmetricBuffer(sym_bytes, dev)
            self.sym = self.sym_buf.tensor
        else:
            self.sym_buf = None
            self.sym = torch.zeros(sym_bytes, dtype=torch.uint8, device=dev)
        self.q_all = self.sym[: q_bytes].view(torch.float32).view(R * B, D)
        self.mine = self.sym[self.mine_off: self.mine_off + m_bytes].view(torch.int64).view(2, R, B, k)  # own queries: [score|idx][shard]
        self.flags = self.sym[self.flags_off: self.flags_off + 2 * R * 8].view(torch.int64)                # [exchange][source rank]
        self.epochs = torch.zeros(2, dtype=torch.int64, device=dev)                                         # local, one per exchange
        self.peers = None          # _lib.Peers of this slot's region on every rank (p2p mode)
        self.loc = torch.empty(2, R * B, k, dtype=torch.int64, device=dev)          # [score bits | global idx] of all R*B queries
        self.gathered = None                                                         # every shard's lists (NCCL mode only)
        n_len64 = (B + 1) // 2
        self.d_final = torch.zeros(2 * B * k + n_len64, dtype=torch.int64, device=dev)
        self.h_final = torch.zeros(2 * B * k + n_len64, dtype=torch.int64).pin_memory()
        self.final_scores = self.d_final[: B * k].view(torch.float64).view(B, k)
        self.final_idx = self.d_final[B * k: 2 * B * k].view(B, k)
        self.final_counts = self.d_final[2 * B * k:].view(torch.int32)[:B]
        self.h_scores = self.h_final[: B * k].view(torch.float64).view(B, k)
        self.h_idx = self.h_final[B * k: 2 * B * k].view(B, k)
        self.h_counts = self.h_final[2 * B * k:].view(torch.int32)[:B].numpy()
        with torch.cuda.device(dev):
            n = e.lib.t2p_retrieve_topk_workspace(R * B, e.db.shape[0], e.db.shape[1], k)
            self.ws_topk = torch.empty(max(256, n), dtype=torch.uint8, device=dev)
            self.done = torch.cuda.Event()
            self.done.record()
        self.graphs = {}


class ShardedOnlineRetrievalEngine:
    """Data-parallel queries over a row-sharded DB: one ``OnlineRetrievalEngine`` per rank over its shard (``idx_base`` =
    first row).  Every rank brings ITS OWN batch of B queries per step (global batch R*B):

        device tokeniser + text encoder (own B queries)
        -> all-gather of the query embeddings            (B*D*4 = 64 KB per rank)
        -> local top-k of all R*B queries against the shard
        -> all-gather of the per-shard top-k lists       (R*B*k*16 bytes per rank)
        -> ``t2p_topk_merge`` of the R lists of the own B queries.

    No rank repeats another rank's text encoding; the exchange volume is independent of the DB size.  ``submit`` /
    ``collect`` pipeline ``engine.depth`` batches like the single-GPU engine (collectives are issued in the same slot
    order on every rank)."""

    def __init__(self, engine: OnlineRetrievalEngine, group=None, exchange: str = "p2p"):
        """``exchange``: "p2p" = own push/wait kernels over CUDA-IPC peer memory (NVLink; CUDA-graph capturable, nothing of
        NCCL on the data path), "nccl" = two ``all_gather_into_tensor`` calls per step."""
        import torch.distributed as dist

        self.dist, self.group, self.eng = dist, group, engine
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if exchange not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        if exchange == "p2p" and self.world > _lib.MAX_PEERS:
            raise ValueError(f"p2p exchange supports up to {_lib.MAX_PEERS} ranks (one NVLink domain)")
        self.exchange = exchange
        self.slots = [_ShardSlot(self) for _ in range(engine.depth)]
        self._inflight = collections.deque()
        self._next = 0
        if exchange == "p2p":
            self._map_peers()
        else:
            # one communicator (hence one NCCL stream) per slot: collectives of different slots must not queue behind
            # each other, or the shared NCCL stream would serialise the batches in flight
            ranks = dist.get_process_group_ranks(group) if group is not None else list(range(self.world))
            self.groups = [group] + [dist.new_group(ranks) for _ in range(engine.depth - 1)]
            R, B, k = self.world, engine.B, engine.k
            for s in self.slots:
                s.gathered = torch.empty(R * 2, R * B, k, dtype=torch.int64, device=engine.device)

    def _map_peers(self):
        """Exchange the CUDA IPC handles of every slot's symmetric region (host side, once) and map the peers' regions."""
        e = self.eng
        torch.cuda.synchronize(e.device)
        everyone = [None] * self.world
        self.dist.all_gather_object(everyone, [s.sym_buf.handle for s in self.slots], group=self.group)
        for i, s in enumerate(self.slots):
            p = _lib.Peers()
            p.n_peers, p.my_rank = self.world, self.rank
            for r in range(self.world):
                p.base[r] = s.sym_buf.ptr if r == self.rank else s.sym_buf.open_peer(everyone[r][i])
            s.peers = p
        self.dist.barrier(group=self.group)  # nobody pushes before every rank has mapped every region

    def close(self):
        """Unmap the peers' regions and free the own ones (collective: call on every rank before the process group dies)."""
        if self.exchange == "p2p":
            torch.cuda.synchronize(self.eng.device)
            self.dist.barrier(group=self.group)
            for s in self.slots:
                s.graphs = {}
                s.q_all = s.mine = s.flags = s.sym = None
                s.sym_buf.close()

    def _topk_all(self, s, db):
        e, R = self.eng, self.world
        _lib.check(
            e.lib.t2p_retrieve_topk_ex(s.q_all.data_ptr(), db.data_ptr(), R * e.B, db.shape[0], db.shape[1], e.k, e.idx_base,
                                       e.db_norm2_max.data_ptr(), e.topk_flags, s.loc[0].data_ptr(), s.loc[1].data_ptr(),
                                       e.stats.data_ptr(), s.ws_topk.data_ptr(), s.ws_topk.numel(), _lib.stream_ptr(e.device)),
            "retrieve_topk",
        )

    def enqueue_exchange(self, db: Optional[torch.Tensor] = None, slot: int = 0):
        """Everything after the text encoder of ``engine.slots[slot]``: gather queries, local top-k, gather lists, merge."""
        e, s, R = self.eng, self.slots[slot], self.world
        es = e.slots[slot]
        B, k, D = e.B, e.k, e.D
        db = e.db if db is None else db
        st = _lib.stream_ptr(e.device)
        if self.exchange == "p2p":
            lib, blk = e.lib, B * k * 8
            # q [B,D] -> every peer's q_all[rank]; then wait for the R blocks of this rank's q_all
            _lib.check(lib.t2p_peer_push(s.peers, es.q.data_ptr(), 0, s.q_all_off + self.rank * B * D * 4, B * D * 4,
                                         None, 0, 0, 0, s.flags_off + (0 * R + self.rank) * 8, s.epochs[0:].data_ptr(), st), "peer_push")
            _lib.check(lib.t2p_peer_wait(s.flags[0:].data_ptr(), R, s.epochs[0:].data_ptr(), st), "peer_wait")
            self._topk_all(s, db)
            # lists of peer j's queries (rows j*B..) -> peer j's mine[score|idx][rank]; then wait for the R lists of the own rows
            _lib.check(lib.t2p_peer_push(s.peers, s.loc[0].data_ptr(), blk, s.mine_off + (0 * R + self.rank) * blk, blk,
                                         s.loc[1].data_ptr(), blk, s.mine_off + (1 * R + self.rank) * blk, blk,
                                         s.flags_off + (1 * R + self.rank) * 8, s.epochs[1:].data_ptr(), st), "peer_push")
            _lib.check(lib.t2p_peer_wait(s.flags[R:].data_ptr(), R, s.epochs[1:].data_ptr(), st), "peer_wait")
        else:
            self.dist.all_gather_into_tensor(s.q_all, es.q, group=self.groups[slot])
            self._topk_all(s, db)
            self.dist.all_gather_into_tensor(s.gathered, s.loc, group=self.groups[slot])
            # rows of this rank's queries from every shard: [shard, {score,idx}, B, k] -> [{score,idx}, shard, B, k]
            s.mine.copy_(s.gathered.view(R, 2, R, B, k)[:, :, self.rank].permute(1, 0, 2, 3))
        _lib.check(
            e.lib.t2p_topk_merge(s.mine[0].data_ptr(), s.mine[1].data_ptr(), R, B, k, k, s.final_scores.data_ptr(),
                                 s.final_idx.data_ptr(), st),
            "topk_merge",
        )

    def enqueue_step(self, db: Optional[torch.Tensor] = None, slot: int = 0, tokenize: bool = True):
        e = self.eng
        if tokenize:
            e.enqueue_tokenize(slot)
        e.enqueue_encode(slot=slot)
        self.enqueue_exchange(db, slot)

    def capture(self, key=0, db: Optional[torch.Tensor] = None, slot: int = 0):
        """p2p mode: capture one whole sharded step of ``slot`` (tokeniser .. merge, pushes and waits included) into a CUDA
        graph.  Collective: every rank must capture (it runs one real step first) and later replay in the same order."""
        if self.exchange != "p2p":
            raise RuntimeError("only the p2p exchange is captured into CUDA graphs")
        e = self.eng
        with torch.cuda.device(e.device):
            st = torch.cuda.Stream()
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                self.enqueue_step(db, slot=slot)
            torch.cuda.current_stream().wait_stream(st)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.enqueue_step(db, slot=slot)
                self.slots[slot].final_counts.copy_(e.slots[slot].lengths)
            self.slots[slot].graphs[key] = g
            e.slots[slot].raw.pop(("sharded", key), None)
        return g

    def capture_all(self, key=0, db: Optional[torch.Tensor] = None):
        for i in range(self.eng.depth):
            self.capture(key, db, slot=i)

    def replay(self, key=0, slot: int = 0):
        self.slots[slot].graphs[key].replay()

    def replay_plan(self, keys: Sequence, first_slot: int = 0):
        return _replay_plan(self.slots, self.eng.slots, self.eng.device, keys, first_slot)

    def replay_many(self, plan, fork_join: bool = True):
        """Collective in p2p mode: every rank must launch the same sequence of (slot, step)."""
        e = self.eng
        _lib.check(e.lib.t2p_serving_replay_many(*plan, _lib.stream_ptr(e.device), int(fork_join)), "serving_replay_many")

    def _enqueue_query(self, slot: int, descriptions, graph_key=None):
        e, s = self.eng, self.slots[slot]
        on_device = e._stage(descriptions, slot)
        if on_device and graph_key is not None and graph_key in s.graphs:
            s.graphs[graph_key].replay()
        else:
            self.enqueue_step(slot=slot, tokenize=on_device)
            s.final_counts.copy_(e.slots[slot].lengths)
        s.h_final.copy_(s.d_final, non_blocking=True)

    def _check(self, s: _ShardSlot):
        self.eng._check_counts(s)  # same checks as the local engine, on this slot's token counts

    def query(self, descriptions: List[str], graph_key=None):
        """This rank's B strings -> (idx [B,k] global int64, scores [B,k] float64) numpy; collective: every rank calls it."""
        if self._inflight:
            raise RuntimeError("query() while submitted batches are in flight: collect() them first")
        e, s = self.eng, self.slots[0]
        st = e.slots[0].stream
        if st is not None:
            with torch.cuda.stream(st):
                self._enqueue_query(0, descriptions, graph_key)
            st.synchronize()
        else:
            self._enqueue_query(0, descriptions, graph_key)
            torch.cuda.current_stream(e.device).synchronize()
        self._check(s)
        return s.h_idx.numpy(), s.h_scores.numpy()

    def submit(self, descriptions: List[str], graph_key=None) -> int:
        e = self.eng
        if e.depth < 2:
            raise RuntimeError("submit() needs an engine built with depth >= 2")
        if len(self._inflight) >= e.depth:
            raise RuntimeError(f"{e.depth} batches already in flight: collect() first")
        slot = self._next
        s, es = self.slots[slot], e.slots[slot]
        g = s.graphs.get(graph_key) if graph_key is not None else None
        if g is None or not _native_submit(e.lib, descriptions, e.B, es, ("sharded", graph_key), g, s.d_final, s.h_final, s.done, es.stream):
            with torch.cuda.stream(es.stream):
                self._enqueue_query(slot, descriptions, graph_key)
                s.done.record()
        self._next = (self._next + 1) % e.depth
        self._inflight.append(slot)
        return slot

    def collect(self):
        slot = self._inflight.popleft()
        s = self.slots[slot]
        s.done.synchronize()
        self._check(s)
        return s.h_idx.numpy(), s.h_scores.numpy()
