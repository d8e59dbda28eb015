"""text2pos_cvpr2022_b200 -- B200-native (sm_100a) implementation of the Text2Pos hot path.

Coarse cell-retrieval forward pass (PointNet++ object encoder, object/cell aggregation, biLSTM text encoder),
all-pairs top-k over the cell database, and the fine SuperGlue attention/Sinkhorn head, behind the reference's
Python module API.  The compute lives in ``csrc/*.cu`` (one C-ABI shared library, ``include/text2pos_b200.h``);
this package is the host-side mirror of the reference interface.  No CPU fallback exists.

Importing the package does not load the native library; the first compute call does (``_lib.load()``).
"""
__version__ = "0.1.0"

from . import synthetic  # noqa: F401
from .runtime import AttrDict, default_args  # noqa: F401

__all__ = ["synthetic", "AttrDict", "default_args"]
