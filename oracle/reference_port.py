"""Oracle (test infrastructure): the reference's CPU execution path, as timed by ``bench.py``'s CPU arms.

``/root/reference`` cannot travel to the GPU box, so ``bench.py --impl reference`` / ``cpu_baseline`` time this
port (``kind: "port"``): the text encoder built from the same library calls the reference makes
(``nn.Embedding`` -> ``pack_padded_sequence`` -> ``nn.LSTM`` -> mean of h_n, ``models/modules.py:59-92``) followed by
the verbatim float64 numpy retrieval loop (``training/coarse.py:134-140``).  PINNED against the reference's
golden vectors in ``tests/test_oracle_golden.py``.
"""
from typing import Dict, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .retrieval import reference_loop
from .text import tokenize


class LanguageEncoderPort(nn.Module):
    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str, known_words: Dict[str, int]):
        super().__init__()
        emb = sd[prefix + "word_embedding.weight"]
        V, D = emb.shape
        self.known_words = known_words
        self.word_embedding = nn.Embedding(V, D, padding_idx=0)
        self.lstm = nn.LSTM(input_size=D, hidden_size=D, bidirectional=True, num_layers=1)
        self.load_state_dict({k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)})
        self.eval()

    @torch.no_grad()
    def forward(self, descriptions: Sequence[str]) -> torch.Tensor:
        tokens, lengths = tokenize(descriptions, self.known_words)
        padded = torch.from_numpy(tokens)
        embedded = self.word_embedding(padded)
        packed = nn.utils.rnn.pack_padded_sequence(embedded, torch.tensor(lengths), batch_first=True, enforce_sorted=False)
        B, D = len(descriptions), self.word_embedding.embedding_dim
        h = torch.zeros(2, B, D)
        c = torch.zeros(2, B, D)
        _, (h, c) = self.lstm(packed, (h, c))
        return torch.mean(h, dim=0)


class CoarseOnlinePort:
    """text -> embedding -> float64 scores -> argsort -> top-k, the reference's eval_epoch query path on the CPU."""

    def __init__(self, sd, known_words, cell_encodings: np.ndarray, k: int):
        self.enc = LanguageEncoderPort(sd, "language_encoder.", known_words)
        self.cells = np.zeros(cell_encodings.shape)  # float64 holder, training/coarse.py:100
        self.cells[:] = cell_encodings
        self.k = k

    def step(self, descriptions: Sequence[str]) -> np.ndarray:
        text = F.normalize(self.enc(descriptions)).cpu().detach().numpy()
        return reference_loop(self.cells, text, self.k)
