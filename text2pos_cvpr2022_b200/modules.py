"""Drop-in for the reference's ``models/modules.py``: ``get_mlp`` and ``LanguageEncoder``.

Same constructor signatures and ``state_dict`` keys (``word_embedding.weight``,
``lstm.{weight,bias}_{ih,hh}_l0[_reverse]``); ``forward`` runs the sm_100a LSTM kernels
(tcgen05 recurrence ``csrc/lstm_tc.cu`` for H = 256, ``csrc/lstm.cu`` otherwise) through ``t2p_lstm_encode`` instead of cuDNN.
"""
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import _lib, packing
from .runtime import PackedModule


def get_mlp(channels: List[int], add_batchnorm: bool = True) -> nn.Sequential:
    """Parameter container with the reference layout (models/modules.py:11-36): one
    ``Sequential(Linear, [BatchNorm1d], ReLU)`` per layer -- note the TRAILING ReLU.  The CUDA path reads the
    parameters through ``packing.pack_mlp_layer`` and never calls this container."""
    blocks = []
    for c_in, c_out in zip(channels[:-1], channels[1:]):
        layer = [nn.Linear(c_in, c_out)]
        if add_batchnorm:
            layer.append(nn.BatchNorm1d(c_out))
        layer.append(nn.ReLU())
        blocks.append(nn.Sequential(*layer))
    return nn.Sequential(*blocks)


def tokenize(descriptions: Sequence[str], known_words: Dict[str, int]):
    """Host tokeniser with the reference's rules (models/modules.py:60-72): strip '.' and ',', lower, split,
    OOV -> 0; zero padded int32 [B, T_max] + lengths [B]."""
    rows = [[known_words.get(w, 0) for w in d.replace(".", "").replace(",", "").lower().split()] for d in descriptions]
    lengths = np.fromiter((len(r) for r in rows), dtype=np.int32, count=len(rows))
    T = int(lengths.max()) if len(rows) else 0
    tokens = np.zeros((len(rows), max(T, 1)), dtype=np.int32)
    for i, r in enumerate(rows):
        tokens[i, : len(r)] = r
    return tokens, lengths


class LanguageEncoder(PackedModule):
    def __init__(self, known_words, embedding_dim, bi_dir, num_layers=1):
        super().__init__()
        if not bi_dir or num_layers != 1:
            raise NotImplementedError("the B200 text encoder implements the reference configuration: 1-layer biLSTM")
        self.known_words = {c: (i + 1) for i, c in enumerate(known_words)}
        self.known_words["<unk>"] = 0
        self.word_embedding = nn.Embedding(len(self.known_words), embedding_dim, padding_idx=0)
        self.lstm = nn.LSTM(input_size=embedding_dim, hidden_size=embedding_dim, bidirectional=True, num_layers=1)

    def _t2p_pack(self, sd):
        bb = packing.BlobBuilder()
        desc = packing.pack_lstm(bb, sd, "")
        return bb.finish(), desc

    def encode_tokens(self, tokens: torch.Tensor, lengths: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        """tokens [B,T] int32, lengths [B] int32, both on the module's device -> [B, D]."""
        weights, desc = self.t2p_packed()
        return lstm_encode(weights, desc, tokens, lengths, normalize, self)

    def encode(self, descriptions: Sequence[str], normalize: bool = False) -> torch.Tensor:
        if len(descriptions) == 0:
            return torch.zeros(0, self.word_embedding.embedding_dim, device=self.device)
        tokens, lengths = tokenize(descriptions, self.known_words)
        if int(lengths.min()) < 1:
            raise ValueError("LanguageEncoder: empty description (the reference's packed LSTM rejects length 0 too)")
        dev = self.t2p_device()
        tok = torch.from_numpy(tokens).pin_memory().to(dev, non_blocking=True)
        ln = torch.from_numpy(lengths).pin_memory().to(dev, non_blocking=True)
        return self.encode_tokens(tok, ln, normalize)

    def forward(self, descriptions):
        """[B, D] = mean of the two final hidden states (NOT normalised), models/modules.py:59-92."""
        return self.encode(descriptions, normalize=False)

    @property
    def device(self):
        return next(self.lstm.parameters()).device


def lstm_encode(weights, desc, tokens, lengths, normalize, owner: PackedModule) -> torch.Tensor:
    lib = _lib.load()
    _lib.require_cuda(tokens, "tokens")
    tokens = tokens.to(torch.int32).contiguous()
    lengths = lengths.to(torch.int32).contiguous()
    B, T = tokens.shape
    H = desc.hidden
    out = torch.empty(B, H, dtype=torch.float32, device=tokens.device)
    with torch.cuda.device(tokens.device):
        ws = owner.t2p_workspace(lib.t2p_lstm_encode_workspace(B, H), tokens.device)
        _lib.check(
            lib.t2p_lstm_encode(weights.handle, desc, _lib.ptr(tokens), _lib.ptr(lengths), B, T, 1 if normalize else 0,
                                _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(tokens.device)),
            "lstm_encode",
        )
    return out
