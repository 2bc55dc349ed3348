import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
MODEL_DIR = os.path.join(GOLDEN, "model_dir")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a GPU (or without the built library) skips the
    `gpu` tests instead of erroring in their fixtures.  On a GPU box nothing is skipped: a missing
    libcia.so there must fail loudly (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (sm_100a); run with gpurun")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_tiny():
    return dict(np.load(os.path.join(GOLDEN, "tiny_field.npz")))


@pytest.fixture(scope="session")
def golden_config1():
    return dict(np.load(os.path.join(GOLDEN, "config1_seed0.npz")))


@pytest.fixture(scope="session")
def model_dir():
    return MODEL_DIR


@pytest.fixture(scope="session")
def artifacts(model_dir):
    from cell_image_analysis_b200.artifacts import load_model_dir
    return load_model_dir(model_dir)


@pytest.fixture(scope="session")
def oracle_weights(artifacts):
    ae = artifacts["autoencoder"]
    return {"kernels": ae["kernels"], "biases": ae["biases"], "bns": ae["bns"]}


@pytest.fixture(scope="session")
def screener(model_dir):
    """The drop-in class on cuda:0 (fp32 CAE path)."""
    from cell_image_analysis_b200.screening import ProductionMutantScreening
    return ProductionMutantScreening(model_dir, segmenter=lambda ch: None, device=0, precision=0)


@pytest.fixture(scope="session")
def field_config1():
    from cell_image_analysis_b200 import synth
    return synth.make_field(0)
