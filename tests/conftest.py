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
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_sail_golden(name):
    """Fixture written by oracle/make_golden.py from the unmodified reference."""
    arr = dict(np.load(os.path.join(GOLDEN, f"sail_{name}.npz")))
    with open(os.path.join(GOLDEN, f"sail_{name}.json")) as f:
        meta = json.load(f)
    params = {k[len("param::"):]: v for k, v in arr.items() if k.startswith("param::")}
    grads = {k[len("grad::"):]: v for k, v in arr.items() if k.startswith("grad::")}
    return arr, meta, params, grads


SAIL_CASES = ["syn", "wd", "wd_clamp", "untied"]


@pytest.fixture(params=SAIL_CASES)
def sail_golden(request):
    return (request.param,) + load_sail_golden(request.param)


def load_ark_golden(name):
    """Decoder-only ARK fixture (oracle/make_golden.py::ark_case) from the unmodified reference."""
    arr = dict(np.load(os.path.join(GOLDEN, f"ark_{name}.npz")))
    with open(os.path.join(GOLDEN, f"ark_{name}.json")) as f:
        meta = json.load(f)
    params = {k[len("param::"):]: v for k, v in arr.items() if k.startswith("param::")}
    grads = {k[len("grad::"):]: v for k, v in arr.items() if k.startswith("grad::")}
    return arr, meta, params, grads


ARK_CASES = ["syn", "wd"]


def load_tsail_golden(name):
    """Transformer KG-VAE fixture (oracle/make_golden.py::tsail_case) from the unmodified reference, dropout 0."""
    arr = dict(np.load(os.path.join(GOLDEN, f"tsail_{name}.npz")))
    with open(os.path.join(GOLDEN, f"tsail_{name}.json")) as f:
        meta = json.load(f)
    params = {k[len("param::"):]: v for k, v in arr.items() if k.startswith("param::")}
    grads = {k[len("grad::"):]: v for k, v in arr.items() if k.startswith("grad::")}
    return arr, meta, params, grads


TSAIL_CASES = ["syn", "wd"]
