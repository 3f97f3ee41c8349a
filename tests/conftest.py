import json
import os
import subprocess
import sys
from types import SimpleNamespace

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vision-sr_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "ref_bfgs.json")) as fh:
        return json.load(fh)


def make_test_data(golden):
    """A DatasetDetails-like record built from the golden file's vocabulary."""
    w2i = dict(golden["word2id"])
    i2w = {v: k for k, v in w2i.items()}
    i2w[3] = "constant"  # model.py:452
    return SimpleNamespace(word2id=w2i, id2word=i2w,
                           total_variables=list(golden["total_variables"]),
                           total_coefficients=[], una_ops=[], bin_ops=[], rewrite_functions=[])


@pytest.fixture(scope="session")
def test_data(golden):
    return make_test_data(golden)


def make_cfg(R, norm="MSE", idx_remove=False, **extra):
    b = SimpleNamespace(n_restarts=R, add_coefficients_if_not_existing=False,
                        idx_remove=idx_remove, normalization_type=norm, stop_time=1e9, **extra)
    return SimpleNamespace(bfgs=b)


@pytest.fixture(scope="session")
def hostsim():
    """ctypes handle on the g++ host simulation of the kernel cores (test-only)."""
    import ctypes
    d = os.path.join(ROOT, "oracle", "hostsim")
    subprocess.check_call(["make", "-s", "-C", d])
    return ctypes.CDLL(os.path.join(d, "libvsr_hostsim.so"))
