"""Compatibility layer for running the reference's unchanged callers (``evaluation/pipeline.py``, the KITTI360Pose data
layer) on this image:

* ``shims/`` (stand-ins for ``easydict`` and the data-layer slice of ``torch_geometric``) goes to the END of ``sys.path``,
  so a real installation of either package always wins;
* numpy-2 removals the reference still imports are aliased: ``np.int`` (``models/modules.py:69``),
  ``numpy.lib.arraysetops`` (``training/utils.py:4``), ``numpy.lib.function_base`` (``dataloading/kitti360pose/poses.py:9``).

``install()`` is idempotent.  ``python -m text2pos_cvpr2022_b200.compat <module> [args...]`` installs the layer and then runs
``<module>`` as ``__main__`` (e.g. ``evaluation.pipeline`` with the reference checkout on ``PYTHONPATH`` after this
repository): see INTEGRATION.md section 1."""
import os
import runpy
import sys
import types

SHIMS_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "shims")


def install() -> None:
    import numpy as np

    if SHIMS_DIR not in sys.path:
        sys.path.append(SHIMS_DIR)
    if not hasattr(np, "int"):
        np.int = int  # noqa: NPY001 -- removed in numpy 1.24; the reference's LanguageEncoder pads with dtype=np.int
    for name, attrs in (("numpy.lib.arraysetops", {"isin": np.isin, "unique": np.unique, "in1d": np.isin}),
                        ("numpy.lib.function_base", {"flip": np.flip})):
        if name in sys.modules:
            continue
        try:
            __import__(name)
        except ImportError:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            m.__doc__ = "alias installed by text2pos_cvpr2022_b200.compat (module removed in numpy 2)"
            sys.modules[name] = m


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m text2pos_cvpr2022_b200.compat <module> [args...]")
    install()
    sys.argv = argv
    runpy.run_module(argv[0], run_name="__main__", alter_sys=True)


if __name__ == "__main__":
    main()
