"""Oracle (test infrastructure): the reference's text encoder.

Restates ``LanguageEncoder.forward`` (``models/modules.py:59-92``) and
``CellRetrievalNetwork.encode_text`` (``models/cell_retrieval.py:69-75``):
tokenise (drop '.' and ',', lower-case, whitespace split, OOV -> 0), embed
(``padding_idx=0`` row is whatever the state_dict holds -- zeros after
construction), run a 1-layer bidirectional LSTM with zero initial state over
each sequence's OWN length (packed sequences), average the two final hidden
states, L2-normalise.

PINNED: checked against the reference ``LanguageEncoder`` imported from
``/root/reference`` (``tests/golden/make_golden.py`` -> ``language_encoder_*.npz``).
Gate order of ``torch.nn.LSTM`` weights: i, f, g, o.
"""
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def tokenize(descriptions: Sequence[str], known_words: Dict[str, int]) -> Tuple[np.ndarray, np.ndarray]:
    """-> (tokens [B, T_max] int64 zero padded, lengths [B] int64).  models/modules.py:60-72."""
    idx = [
        [known_words.get(w, 0) for w in d.replace(".", "").replace(",", "").lower().split()]
        for d in descriptions
    ]
    lengths = np.array([len(w) for w in idx], dtype=np.int64)
    tokens = np.zeros((len(idx), int(lengths.max()) if len(idx) else 0), dtype=np.int64)
    for i, w in enumerate(idx):
        tokens[i, : len(w)] = w
    return tokens, lengths


def _lstm_direction(x, lengths, w_ih, w_hh, b_ih, b_hh, reverse: bool) -> torch.Tensor:
    """x [B,T,D] -> final hidden state [B,H] of one direction, each row over its own length."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = torch.zeros(B, H, dtype=x.dtype)
    c = torch.zeros(B, H, dtype=x.dtype)
    for step in range(T):
        # forward: t = step; backward: row b visits t = len_b-1-step (packed-sequence semantics)
        t = (lengths - 1 - step) if reverse else torch.full_like(lengths, step)
        active = (step < lengths)
        xt = x[torch.arange(B), t.clamp(min=0)]
        gates = xt @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
        i, f, g, o = gates.chunk(4, dim=1)
        c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h_new = torch.sigmoid(o) * torch.tanh(c_new)
        m = active[:, None]
        c = torch.where(m, c_new, c)
        h = torch.where(m, h_new, h)
    return h


def language_encoder(sd: Dict[str, torch.Tensor], prefix: str, tokens: np.ndarray, lengths: np.ndarray) -> torch.Tensor:
    """-> mean of the forward/backward final hidden states, [B, D] (NOT normalised)."""
    tok = torch.as_tensor(tokens, dtype=torch.long)
    ln = torch.as_tensor(lengths, dtype=torch.long)
    x = sd[prefix + "word_embedding.weight"][tok]  # [B,T,D]
    hs = []
    for suffix, rev in (("", False), ("_reverse", True)):
        hs.append(
            _lstm_direction(
                x, ln,
                sd[f"{prefix}lstm.weight_ih_l0{suffix}"], sd[f"{prefix}lstm.weight_hh_l0{suffix}"],
                sd[f"{prefix}lstm.bias_ih_l0{suffix}"], sd[f"{prefix}lstm.bias_hh_l0{suffix}"], rev,
            )
        )
    return torch.stack(hs).mean(dim=0)  # models/modules.py:90


def encode_text(sd, descriptions: Sequence[str], known_words: Dict[str, int]) -> torch.Tensor:
    """``CellRetrievalNetwork.encode_text``: [B, D] unit-norm rows."""
    tokens, lengths = tokenize(descriptions, known_words)
    return F.normalize(language_encoder(sd, "language_encoder.", tokens, lengths))
