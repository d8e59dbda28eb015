"""ctypes binding of the C ABI declared in ``include/text2pos_b200.h``.

There is NO fallback: if ``lib/libtext2pos_b200.so`` is missing or a call fails, a
``RuntimeError`` is raised.  Build the library with ``python -m text2pos_cvpr2022_b200.build``
(or ``__graft_entry__.build()``).
"""
import ctypes as C
import os
import threading

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libtext2pos_b200.so")

MAX_NEIGHBORS = 32
KNN_K = 8
MAX_GNN_LAYERS = 32

T2P_OK = 0
RETRIEVE_FORCE_GENERIC = 1
RETRIEVE_FORCE_RESCAN = 2


def retrieve_max_ctas(n: int) -> int:
    """``T2P_RETRIEVE_MAX_CTAS(n)`` flag bits."""
    return (int(n) & 0xff) << 8


class LinearDesc(C.Structure):
    _fields_ = [("w_off", C.c_int64), ("b_off", C.c_int64), ("k", C.c_int32), ("n", C.c_int32)]


class PointNet2Desc(C.Structure):
    _fields_ = [
        ("sa_l1", LinearDesc * 3),
        ("sa_l2", LinearDesc * 3),
        ("sa_radius_sq", C.c_float * 3),
        ("ga_l1", LinearDesc),
        ("ga_l2", LinearDesc),
        ("lin1", LinearDesc),
        ("lin2", LinearDesc),
        ("self_loop_quirk", C.c_int32),
        ("reserved", C.c_int32),
        ("ga_l2_tc_off", C.c_int64),
        ("sa_l2_tc_off", C.c_int64 * 3),
        ("dense_tc_off", C.c_int64 * 6),
    ]


class ObjEncDesc(C.Structure):
    _fields_ = [
        ("mlp_pointnet", LinearDesc),
        ("color_l1", LinearDesc),
        ("color_l2", LinearDesc),
        ("pos_l1", LinearDesc),
        ("pos_l2", LinearDesc),
        ("merge", LinearDesc),
        ("embed_dim", C.c_int32),
    ]


class CellAggDesc(C.Structure):
    _fields_ = [
        ("edge_ab", LinearDesc),
        ("edge_l2", LinearDesc),
        ("lin_l1", LinearDesc),
        ("lin_l2", LinearDesc),
        ("embed_dim", C.c_int32),
    ]


class LstmDesc(C.Structure):
    _fields_ = [("xproj_off", C.c_int64), ("whh_off", C.c_int64), ("whh_reg_off", C.c_int64), ("xproj4_off", C.c_int64),
                ("whh_tc_off", C.c_int64), ("vocab", C.c_int32), ("hidden", C.c_int32), ("path", C.c_int32),
                ("max_groups", C.c_int32)]


MAX_PEERS = 8


class Peers(C.Structure):
    _fields_ = [("base", C.c_void_p * MAX_PEERS), ("n_peers", C.c_int32), ("my_rank", C.c_int32)]


class SuperGlueDesc(C.Structure):
    _fields_ = [
        ("q", LinearDesc * MAX_GNN_LAYERS),
        ("k", LinearDesc * MAX_GNN_LAYERS),
        ("v", LinearDesc * MAX_GNN_LAYERS),
        ("merge", LinearDesc * MAX_GNN_LAYERS),
        ("mlp0", LinearDesc * MAX_GNN_LAYERS),
        ("mlp3", LinearDesc * MAX_GNN_LAYERS),
        ("is_cross", C.c_int32 * MAX_GNN_LAYERS),
        ("final_proj", LinearDesc),
        ("num_gnn_layers", C.c_int32),
        ("dim", C.c_int32),
        ("sinkhorn_iters", C.c_int32),
        ("bin_score", C.c_float),
        ("match_threshold", C.c_float),
        ("tc_w_off", C.c_int64),
        ("tc_b_off", C.c_int64),
    ]


_P = C.c_void_p
_I = C.c_int
_SZ = C.c_size_t

# name -> (restype, argtypes).  Every symbol of include/text2pos_b200.h is listed (tests check the export list).
PROTOTYPES = {
    "t2p_version": (_I, []),
    "t2p_last_error": (C.c_char_p, []),
    "t2p_device_info": (_I, [C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "t2p_weights_create": (_I, [_P, _SZ, C.POINTER(_P)]),
    "t2p_weights_destroy": (_I, [_P]),
    "t2p_weights_device_ptr": (_P, [_P]),
    "t2p_retrieve_topk_workspace": (_SZ, [_I, _I, _I, _I]),
    "t2p_retrieve_topk": (_I, [_P, _P, _I, _I, _I, _I, C.c_int64, _P, _P, _P, _SZ, _P]),
    "t2p_retrieve_topk_ex": (_I, [_P, _P, _I, _I, _I, _I, C.c_int64, _P, _I, _P, _P, _P, _P, _SZ, _P]),
    "t2p_db_row_norm2_max": (_I, [_P, _I, _I, _P, _P]),
    "t2p_topk_merge": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "t2p_enable_peer_access": (_I, [_I]),
    "t2p_ipc_alloc": (_I, [_SZ, C.POINTER(_P), _P]),
    "t2p_ipc_open": (_I, [_P, C.POINTER(_P)]),
    "t2p_ipc_close": (_I, [_P]),
    "t2p_ipc_free": (_I, [_P]),
    "t2p_peer_push": (_I, [C.POINTER(Peers), _P, _SZ, _SZ, _SZ, _P, _SZ, _SZ, _SZ, _SZ, _P, _P]),
    "t2p_peer_wait": (_I, [_P, _I, _P, _P]),
    "t2p_fps": (_I, [_P, _I, _I, _I, _P, _P]),
    "t2p_ball_query": (_I, [_P, _P, _I, _I, _I, C.c_float, _I, _P, _P, _P]),
    "t2p_linear": (_I, [_P, C.POINTER(LinearDesc), _P, _I, _I, _I, _P, _I, _P]),
    "t2p_l2_normalize_rows": (_I, [_P, _I, _I, _I, _P]),
    "t2p_knn_cells": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "t2p_pointnet2_workspace": (_SZ, [C.POINTER(PointNet2Desc), _I, _I]),
    "t2p_pointnet2_forward": (
        _I,
        [_P, C.POINTER(PointNet2Desc), _P, _P, _P, _I, _I, _P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), _P, _SZ, _P],
    ),
    "t2p_object_embed_workspace": (_SZ, [C.POINTER(ObjEncDesc), _I]),
    "t2p_object_embed": (_I, [_P, C.POINTER(ObjEncDesc), _P, _P, _P, _I, _P, _P, _SZ, _P]),
    "t2p_cell_aggregate_workspace": (_SZ, [C.POINTER(CellAggDesc), _I, _I]),
    "t2p_cell_aggregate": (_I, [_P, C.POINTER(CellAggDesc), _P, _P, _I, _I, _I, _P, _P, _P, _SZ, _P]),
    "t2p_vocab_create": (_I, [C.POINTER(C.c_char_p), C.POINTER(C.c_int32), _I, C.POINTER(_P)]),
    "t2p_vocab_destroy": (_I, [_P]),
    "t2p_tokenize": (_I, [_P, C.c_char_p, _SZ, _I, _I, _P, _P, C.POINTER(C.c_int32)]),
    "t2p_vocab_to_device": (_I, [_P]),
    "t2p_stage_texts_capacity": (_SZ, [_I, _SZ]),
    "t2p_stage_texts": (_I, [C.c_char_p, _SZ, _I, _P, _SZ, _P, _P, C.POINTER(_SZ), C.POINTER(_I)]),
    "t2p_tokenize_device": (_I, [_P, _P, _I, _I, _P, _P, _P]),
    "t2p_serving_submit": (_I, [C.c_char_p, _SZ, _I, _P, _SZ, _P, _P, _P, _P, _SZ, _P, _P, C.POINTER(_SZ), C.POINTER(_I)]),
    "t2p_serving_replay_many": (_I, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _I, _P, _I]),
    "t2p_lstm_encode_workspace": (_SZ, [_I, _I]),
    "t2p_lstm_encode": (_I, [_P, C.POINTER(LstmDesc), _P, _P, _I, _I, _I, _P, _P, _SZ, _P]),
    "t2p_superglue_workspace": (_SZ, [_I, _I, _I, _I]),
    "t2p_superglue_forward": (
        _I,
        [_P, C.POINTER(SuperGlueDesc), _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _SZ, _P],
    ),
    "t2p_superglue_forward_gather": (
        _I,
        [_P, C.POINTER(SuperGlueDesc), _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _SZ, _P],
    ),
    "t2p_fixed_points_index": (C.c_uint32, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]),
    "t2p_batch_object_points": (_I, [_P, _P, _P, _I, _I, _P, C.c_uint64, C.c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "t2p_pose_head": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "t2p_pose_accuracy": (_I, [_P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, C.POINTER(C.c_int32), _I, C.POINTER(C.c_double), _I, _P, _P]),
}

_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    """Load the shared library (no GPU needed for loading)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"text2pos_b200: native library {LIB_PATH} is missing; build it with "
                "`python -m text2pos_cvpr2022_b200.build` (there is no CPU / PyTorch fallback)"
            )
        from . import build as _build

        if _build.built_hash() != _build.source_hash():
            raise RuntimeError(
                f"text2pos_b200: {LIB_PATH} was built from different sources than the ones in csrc/ (BUILD_STAMP mismatch); "
                "rebuild with `python -m text2pos_cvpr2022_b200.build`"
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int, what: str = "") -> None:
    if rc != T2P_OK:
        msg = load().t2p_last_error()
        raise RuntimeError(f"text2pos_b200 {what} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr(t) -> int:
    """Device (or host) pointer of a contiguous tensor; None -> NULL."""
    if t is None:
        return None
    assert t.is_contiguous(), "text2pos_b200: tensors passed to the C ABI must be contiguous"
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"text2pos_b200: {what} must live on a CUDA device (sm_100a); there is no CPU path")


class _RawCuda:
    """``__cuda_array_interface__`` view of raw device memory (for ``torch.as_tensor``)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class SymmetricBuffer:
    """A zero-filled cudaMalloc block of this rank that the peers map over CUDA IPC (``t2p_ipc_*``); ``tensor`` is a
    uint8 torch view of the local block, ``handle`` the 64 bytes to ship to the peers, ``open_peer`` maps theirs."""

    def __init__(self, nbytes: int, device):
        self.lib = load()
        self.device = torch.device(device)
        self.nbytes = int(nbytes)
        p, h = _P(), C.create_string_buffer(64)
        with torch.cuda.device(self.device):
            check(self.lib.t2p_ipc_alloc(self.nbytes, C.byref(p), h), "ipc_alloc")
            self.ptr = int(p.value)
            self.tensor = torch.as_tensor(_RawCuda(self.ptr, self.nbytes), device=self.device)
        self.handle = bytes(h.raw)
        self.peer_ptrs = []

    def open_peer(self, handle: bytes) -> int:
        p = _P()
        with torch.cuda.device(self.device):
            check(self.lib.t2p_ipc_open(C.create_string_buffer(handle, 64), C.byref(p)), "ipc_open")
        self.peer_ptrs.append(int(p.value))
        return int(p.value)

    def close(self):
        try:
            with torch.cuda.device(self.device):
                for q in self.peer_ptrs:
                    self.lib.t2p_ipc_close(q)
                self.peer_ptrs = []
                if self.ptr:
                    self.tensor = None
                    self.lib.t2p_ipc_free(self.ptr)
                    self.ptr = 0
        except Exception:
            pass


class Workspace:
    """Grow-only scratch buffer per (device, stream-agnostic) owner."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes: int, device) -> torch.Tensor:
        nbytes = max(int(nbytes), 256)
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != torch.device(device):
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self.buf


class Weights:
    """Owns a ``t2p_weights`` handle (an immutable device copy of a packed float32 blob)."""

    def __init__(self, blob: torch.Tensor, device):
        lib = load()
        blob = blob.detach().to("cpu", torch.float32).contiguous()
        self.n = blob.numel()
        self.device = torch.device(device)
        h = _P()
        with torch.cuda.device(self.device):
            check(lib.t2p_weights_create(blob.data_ptr(), self.n, C.byref(h)), "weights_create")
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                load().t2p_weights_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class Vocab:
    """Owns a ``t2p_vocab`` handle: the native tokeniser of the text encoder (``csrc/tokenize.cu``)."""

    def __init__(self, known_words: dict):
        lib = load()
        words = [w.encode("utf-8") for w in known_words.keys()]
        ids = list(known_words.values())
        n = len(words)
        arr_w = (C.c_char_p * n)(*words)
        arr_i = (C.c_int32 * n)(*ids)
        h = _P()
        check(lib.t2p_vocab_create(arr_w, arr_i, n, C.byref(h)), "vocab_create")
        self.handle = h

    def tokenize_into(self, descriptions, h_tokens: torch.Tensor, h_lengths: torch.Tensor) -> int:
        """Tokenise ASCII ``descriptions`` into the (pinned) int32 host tensors; returns the longest row length."""
        lib = load()
        n = len(descriptions)
        assert h_tokens.dtype == torch.int32 and h_tokens.is_contiguous() and h_tokens.shape[0] >= n and h_lengths.numel() >= n
        blob = ("\0".join(descriptions) + "\0").encode("utf-8")
        longest = C.c_int32(0)
        check(lib.t2p_tokenize(self.handle, blob, len(blob), n, h_tokens.shape[1], h_tokens.data_ptr(), h_lengths.data_ptr(),
                               C.byref(longest)), "tokenize")
        return int(longest.value)

    def to_device(self, device) -> None:
        """Upload the hash table for the device tokeniser (once; not during stream capture)."""
        with torch.cuda.device(device):
            check(load().t2p_vocab_to_device(self.handle), "vocab_to_device")

    def stage_texts(self, descriptions, h_stage: torch.Tensor, d_stage: torch.Tensor = None):
        """Lay ``descriptions`` out in the pinned uint8 staging tensor for ``tokenize_device`` and (ASCII batches, if
        ``d_stage`` is given) enqueue its H2D copy on the current stream; returns (bytes the device needs, all_ascii)."""
        blob = ("\0".join(descriptions) + "\0").encode("utf-8")
        used, ascii_ = _SZ(0), _I(0)
        check(load().t2p_stage_texts(blob, len(blob), len(descriptions), h_stage.data_ptr(), h_stage.numel(),
                                     None if d_stage is None else d_stage.data_ptr(),
                                     None if d_stage is None else stream_ptr(d_stage.device), C.byref(used), C.byref(ascii_)),
              "stage_texts")
        return int(used.value), bool(ascii_.value)

    def tokenize_device(self, d_stage: torch.Tensor, n_texts: int, d_tokens: torch.Tensor, d_lengths: torch.Tensor) -> None:
        check(load().t2p_tokenize_device(self.handle, d_stage.data_ptr(), n_texts, d_tokens.shape[1], d_tokens.data_ptr(),
                                         d_lengths.data_ptr(), stream_ptr(d_tokens.device)), "tokenize_device")

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                load().t2p_vocab_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
