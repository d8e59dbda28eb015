"""Oracle (test infrastructure): ``get_mlp`` blocks evaluated from a state_dict.

Follows ``models/modules.py:11-36`` of the reference: every layer of ``get_mlp``
is ``Sequential(Linear, BatchNorm1d, ReLU)`` (or ``Sequential(Linear, ReLU)``
without batch-norm) and the stack ENDS in a ReLU.  BatchNorm is evaluated in
eval mode (running statistics, eps=1e-5), un-folded, exactly as
``torch.nn.BatchNorm1d.eval()`` does.
"""
from typing import Dict, List

import torch

BN_EPS = 1e-5


def linear(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor) -> torch.Tensor:
    return x @ sd[prefix + "weight"].t() + sd[prefix + "bias"]


def batchnorm_eval(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor) -> torch.Tensor:
    """BatchNorm1d in eval mode over the channel (last) dim of ``x[..., C]``."""
    rm, rv = sd[prefix + "running_mean"], sd[prefix + "running_var"]
    return (x - rm) / torch.sqrt(rv + BN_EPS) * sd[prefix + "weight"] + sd[prefix + "bias"]


def mlp_num_layers(sd: Dict[str, torch.Tensor], prefix: str) -> int:
    n = 0
    while f"{prefix}{n}.0.weight" in sd:
        n += 1
    return n


def get_mlp(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor) -> torch.Tensor:
    """Apply the ``get_mlp`` stack stored under ``prefix`` (e.g. ``"lin."``) to ``x[..., C_in]``."""
    n = mlp_num_layers(sd, prefix)
    assert n > 0, f"no get_mlp under {prefix!r}"
    for i in range(n):
        x = linear(sd, f"{prefix}{i}.0.", x)
        if f"{prefix}{i}.1.running_mean" in sd:
            x = batchnorm_eval(sd, f"{prefix}{i}.1.", x)
        x = torch.relu(x)
    return x


def mlp_channels(sd: Dict[str, torch.Tensor], prefix: str) -> List[int]:
    n = mlp_num_layers(sd, prefix)
    ch = [sd[f"{prefix}0.0.weight"].shape[1]]
    for i in range(n):
        ch.append(sd[f"{prefix}{i}.0.weight"].shape[0])
    return ch
