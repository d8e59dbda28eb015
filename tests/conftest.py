import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200); run with -m gpu")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return z, meta


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library (loads without a GPU; compute calls need one)."""
    from text2pos_cvpr2022_b200 import _lib

    return _lib.load()


def _cuda_ok():
    import torch

    return torch.cuda.is_available()


@pytest.fixture(scope="session")
def coarse_model():
    """Random-weight coarse model (BN statistics randomised so that folding is exercised), eval mode, on cuda:0."""
    import torch

    from text2pos_cvpr2022_b200 import default_args, synthetic as syn
    from text2pos_cvpr2022_b200.cell_retrieval import CellRetrievalNetwork

    m = CellRetrievalNetwork(syn.KNOWN_CLASSES, syn.COLOR_NAMES, syn.known_words(), default_args(embed_dim=256))
    syn.randomize_module_(m, 5, gain=2.0)
    m.eval()
    return m.to("cuda") if _cuda_ok() else m


@pytest.fixture(scope="session")
def fine_model():
    import torch

    from text2pos_cvpr2022_b200 import default_args, synthetic as syn
    from text2pos_cvpr2022_b200.superglue_matcher import SuperGlueMatch

    m = SuperGlueMatch(syn.KNOWN_CLASSES, syn.COLOR_NAMES, syn.known_words(), default_args(embed_dim=128, num_layers=6))
    sd = syn.synth_state_dict([(k, tuple(v.shape)) for k, v in m.state_dict().items()], 7, gain=0.4)
    syn.superglue_peaky_(sd, "superglue.", scale=5.0)
    m.load_state_dict(sd)
    m.eval()
    return m.to("cuda") if _cuda_ok() else m


def cpu_state_dict(module):
    return {k: v.detach().cpu() for k, v in module.state_dict().items()}
