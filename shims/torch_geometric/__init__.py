"""Stand-in for the slice of ``torch_geometric`` that the reference's DATA LAYER imports (``requirements.txt:11``; the
package is not installed in this image and has no wheel offline).

Found only when the real package is absent (``text2pos_cvpr2022_b200.compat.install()`` appends ``shims/`` to the END of
``sys.path``).  Provided: ``torch_geometric.data.{Data, Batch}`` and ``torch_geometric.transforms.{Compose, FixedPoints,
NormalizeScale, Center, RandomRotate}`` -- what ``evaluation/pipeline.py:29,290-293``, ``dataloading/kitti360pose/utils.py:7,
99-109``, ``dataloading/kitti360pose/{cells,poses,eval}.py`` and ``training/utils.py:8`` need.  ``torch_geometric.nn`` is
deliberately NOT provided: the B200 modules replace every model that used it (fps / radius / PointConv / DynamicEdgeConv
run as sm_100a kernels), so nothing on the hot path imports it.  Semantics restated from the published PyG behaviour
([PyG-recalled] in SURVEY.md: parity unpinned)."""
from . import data, transforms  # noqa: F401

__version__ = "0.0-t2p-shim"
