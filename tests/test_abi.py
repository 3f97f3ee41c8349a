"""The C-ABI library loads and exports every symbol include/vsr.h declares (no GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from src.visymre.engine import native


@pytest.fixture(scope="module")
def lib():
    native.build()
    return native.load()


def test_every_declared_symbol_is_exported(lib):
    header = open(os.path.join(ROOT, "include", "vsr.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(vsr_[a-z_0-9]+)\s*\(", header))
    assert declared == set(native.SIGNATURES), declared ^ set(native.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)


def test_defaults_are_scipys(lib):
    o = native.FitOpts()
    lib.vsr_fit_opts_default(ctypes.byref(o))
    assert (o.gtol, o.c1, o.c2, o.xrtol) == (1e-5, 1e-4, 0.9, 0.0)
    assert o.fd_eps == 1.4901161193847656e-08 and o.penalty == 1e6 and o.maxiter_per_k == 200
    assert o.stop_time == 1e9 and o.loss_scale == 1.0


def test_fails_loudly_without_a_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = lib.vsr_create(0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"no CPU path" in lib.vsr_last_error(None)
    from src.visymre.engine import fitter
    with pytest.raises(native.VsrError):
        fitter.Engine()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vision-sr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "hostsim" not in text or f in ("vsr_interp.h", "vsr_bfgs.h", "vsr_isa.h"), f
