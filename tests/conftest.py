import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def pkg():
    """the product package (through its import alias); builds nothing -- the .so must already exist"""
    import mwa_b200
    return mwa_b200


@pytest.fixture(scope="session")
def lib(pkg):
    return pkg._abi.load()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return {k: np.load(os.path.join(GOLDEN, k + ".npz")) for k in ("attention", "gdn", "rounding", "wrapper", "pyramid", "model_rgb")}


@pytest.fixture(scope="session")
def model_keys():
    """state-dict key -> shape of the reference's AutoEncoderRGB_Journal.AutoEncoder (written by oracle/make_golden.py)"""
    import json
    with open(os.path.join(GOLDEN, "model_rgb_keys.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda:0")
