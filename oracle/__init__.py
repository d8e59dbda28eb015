"""CPU oracle for the Text2Pos hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (numpy float32/float64 + CPU torch), the
algorithm of the reference's coarse cell-retrieval forward pass, its all-pairs
top-k and the fine SuperGlue head.  It exists to *check* the CUDA path; it is
never the thing that is shipped or measured.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product package
(``text2pos_cvpr2022_b200``) never imports ``oracle`` and has no CPU fallback.

Parity status (see DESIGN.md "Oracle"):

* PINNED against outputs of the reference itself, generated in the dev
  container by importing ``/root/reference`` (script
  ``tests/golden/make_golden.py``, fixtures ``tests/golden/*.npz``):
    - ``oracle.superglue``   <- ``models/superglue.py:90-330``
    - ``oracle.text``        <- ``models/modules.py:40-96`` (LanguageEncoder)
    - ``oracle.mlp.get_mlp`` <- ``models/modules.py:11-36``
    - ``oracle.retrieval``   <- ``training/coarse.py:134-148``
* PARITY UNPINNED for the pieces whose arithmetic lives in third-party
  packages that are absent from ``/root/reference`` and from this image
  (``torch_geometric`` / ``torch_cluster`` / ``torch_scatter``, unpinned in
  ``requirements.txt:11``): farthest point sampling, ball query (``radius``),
  ``PointConv``, ``DynamicEdgeConv`` and ``global_max_pool``.  Their published
  semantics are restated in ``oracle.pointnet`` / ``oracle.cells`` with every
  tie-break made explicit; the reference ships no test or golden vector for
  them, so they are anchored on the reference's call sites only
  (``models/pointcloud/pointnet2.py:25-37,45-49,80-100``,
  ``models/object_encoder.py:61-142``, ``models/cell_retrieval.py:77-107``).
"""

from . import mlp, pointnet, cells, text, retrieval, superglue, models  # noqa: F401
