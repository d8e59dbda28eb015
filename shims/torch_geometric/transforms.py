"""``torch_geometric.transforms`` stand-ins used by the reference (``evaluation/pipeline.py:290-293``,
``training/coarse.py:188-199``): ``Compose``, ``FixedPoints`` (numpy global RNG, sampling WITH replacement -- PyG's default
``replace=True``), ``Center``, ``NormalizeScale`` (centre on the mean, scale by ``0.999999 / max|pos|``), ``RandomRotate``.
[PyG-recalled]: restated from the published behaviour, parity unpinned."""
import math
import re

import numpy as np
import torch


class Compose:
    def __init__(self, transforms):
        self.transforms = list(transforms)

    def __call__(self, data):
        for t in self.transforms:
            data = t(data)
        return data

    def __repr__(self):
        return "Compose([" + ", ".join(repr(t) for t in self.transforms) + "])"


class FixedPoints:
    def __init__(self, num, replace=True, allow_duplicates=False):
        self.num, self.replace, self.allow_duplicates = int(num), replace, allow_duplicates

    def __call__(self, data):
        n = data.num_nodes
        if self.replace:
            choice = torch.from_numpy(np.random.choice(n, self.num, replace=True)).long()
        elif not self.allow_duplicates:
            choice = torch.randperm(n)[: self.num]
        else:
            choice = torch.cat([torch.randperm(n) for _ in range(math.ceil(self.num / n))])[: self.num]
        for k, v in list(data):
            if bool(re.search("edge", k)):
                continue
            if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n:
                data[k] = v[choice]
        return data

    def __repr__(self):
        return f"FixedPoints({self.num}, replace={self.replace})"


class Center:
    def __call__(self, data):
        data.pos = data.pos - data.pos.mean(dim=-2, keepdim=True)
        return data


class NormalizeScale:
    def __init__(self):
        self.center = Center()

    def __call__(self, data):
        data = self.center(data)
        scale = (1 / data.pos.abs().max()) * 0.999999
        data.pos = data.pos * scale
        return data

    def __repr__(self):
        return "NormalizeScale()"


class RandomRotate:
    """Rotation by a uniform angle in ``degrees`` around ``axis`` (training-time augmentation; not on the inference path)."""

    def __init__(self, degrees, axis=0):
        if isinstance(degrees, (int, float)):
            degrees = (-abs(degrees), abs(degrees))
        self.degrees, self.axis = degrees, axis

    def __call__(self, data):
        a = math.pi * float(np.random.uniform(*self.degrees)) / 180.0
        s, c = math.sin(a), math.cos(a)
        if self.axis == 0:
            m = [[1, 0, 0], [0, c, s], [0, -s, c]]
        elif self.axis == 1:
            m = [[c, 0, -s], [0, 1, 0], [s, 0, c]]
        else:
            m = [[c, s, 0], [-s, c, 0], [0, 0, 1]]
        data.pos = data.pos @ torch.tensor(m, dtype=data.pos.dtype).t()
        return data
