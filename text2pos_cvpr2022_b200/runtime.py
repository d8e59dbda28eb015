"""Host-side plumbing shared by the drop-in modules: packed-weight caching, attribute dicts, argument defaults."""
import warnings
from types import SimpleNamespace
from typing import Any, Dict

import torch
import torch.nn as nn

from . import _lib


class AttrDict(dict):
    """Stand-in for ``easydict.EasyDict`` (not installed here): key AND attribute access, key iteration."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __contains__(self, k):
        return dict.__contains__(self, k)


def arg(args: Any, name: str, default=None):
    """Read a field of an argparse Namespace / EasyDict / dict the way the reference does (``"x" in args``)."""
    if args is None:
        return default
    if isinstance(args, dict):
        return args.get(name, default)
    return getattr(args, name, default)


def default_args(**overrides) -> SimpleNamespace:
    """The fields of ``training/args.py`` that the hot path reads, with the README's coarse settings."""
    d = dict(
        embed_dim=256, num_layers=6, use_features=["class", "color", "position"], variation=0,
        sinkhorn_iters=50, num_mentioned=6, pad_size=16, pointnet_layers=3, pointnet_variation=0,
        pointnet_numpoints=256, pointnet_path=None, pointnet_freeze=False, pointnet_features=2,
        class_embed=False, color_embed=False, top_k=[1, 5, 10], ranking_loss="pairwise", batch_size=64,
    )
    d.update(overrides)
    return SimpleNamespace(**d)


class PackedModule(nn.Module):
    """nn.Module whose parameters are mirrored into a packed, BN-folded device blob for the CUDA kernels.

    The blob is rebuilt lazily when the module moved, a state_dict was loaded, or any parameter/buffer was
    modified in place (detected through tensor version counters).  Inference semantics only: BatchNorm always
    uses its running statistics (eval mode), whatever ``self.training`` says.
    """

    def __init__(self):
        super().__init__()
        self.__dict__["_t2p_cache"] = None
        self.__dict__["_t2p_ws"] = _lib.Workspace()
        self.__dict__["_t2p_warned_train"] = False
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._t2p_invalidate())

    def _t2p_invalidate(self):
        self.__dict__["_t2p_cache"] = None

    def _apply(self, fn, *a, **kw):
        r = super()._apply(fn, *a, **kw)
        self._t2p_invalidate()
        return r

    def __getstate__(self):  # pickling (torch.save(model)) must not try to pickle ctypes handles
        st = self.__dict__.copy()
        st["_t2p_cache"] = None
        st["_t2p_ws"] = None
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self.__dict__["_t2p_cache"] = None
        self.__dict__["_t2p_ws"] = _lib.Workspace()

    def _t2p_signature(self):
        sig = 0
        dev = None
        for t in list(self.parameters()) + list(self.buffers()):
            sig += t._version
            dev = t.device
        return sig, dev

    def _t2p_pack(self, sd: Dict[str, torch.Tensor]):
        """-> (blob tensor (cpu, float32), descriptor-or-dict).  Implemented by subclasses."""
        raise NotImplementedError

    def t2p_device(self) -> torch.device:
        """Device of the parameters; raises loudly unless it is a CUDA device (no CPU fallback exists)."""
        _, dev = self._t2p_signature()
        if dev is None or dev.type != "cuda":
            raise RuntimeError(
                f"{type(self).__name__}: parameters are on {dev}; the B200 path needs a CUDA device "
                "(move the module with .to('cuda')); there is no CPU fallback"
            )
        return dev

    def t2p_packed(self):
        """-> (Weights handle, descriptors) for the device the parameters live on."""
        sig, dev = self._t2p_signature()
        self.t2p_device()
        if self.training and not self._t2p_warned_train:
            self.__dict__["_t2p_warned_train"] = True
            warnings.warn(
                f"{type(self).__name__} is in train() mode; the B200 kernels implement inference (eval-mode "
                "BatchNorm, running statistics) and ignore the flag"
            )
        c = self._t2p_cache
        if c is None or c[0] != (sig, dev):
            blob, desc = self._t2p_pack({k: v for k, v in self.state_dict().items()})
            self.__dict__["_t2p_cache"] = ((sig, dev), _lib.Weights(blob, dev), desc)
            c = self._t2p_cache
        return c[1], c[2]

    def t2p_workspace(self, nbytes: int, device) -> torch.Tensor:
        return self._t2p_ws.get(nbytes, device)
