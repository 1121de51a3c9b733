"""The C-ABI library loads and exports every symbol include/ttlda.h declares (no compute; runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from tuotuo_b200 import build

    return build.build()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ttlda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ttlda_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    syms = header_symbols()
    for must in ["ttlda_dirichlet_expectation_f64", "ttlda_bow_to_csr", "ttlda_csr_to_csc", "ttlda_estep_f64",
                 "ttlda_sstats_csc_f64", "ttlda_sstats_atomic_f64", "ttlda_mstep_f64", "ttlda_elbo_tokens_f64", "ttlda_hyper_stats_f64",
                 "ttlda_theta_refresh_f64", "ttlda_beta_refresh_f64", "ttlda_version", "ttlda_last_error"]:
        assert must in syms


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for s in header_symbols():
        assert hasattr(lib, s), f"{s} declared in ttlda.h but not exported by libttlda.so"


def test_python_binding_matches_header(lib_path):
    from tuotuo_b200 import _native as nv

    assert sorted(nv.SIGNATURES) == header_symbols()
    lib = nv.load()
    assert lib.ttlda_version() == (0 << 16) | (1 << 8) | 0
    assert lib.ttlda_ld_for(100) == 100 and lib.ttlda_ld_for(5) == 8 and lib.ttlda_ld_for(1000) == 1000
    assert nv.ld_for(501) == 504
    assert lib.ttlda_reduce_workspace_bytes() > 0
    assert lib.ttlda_mstep_workspace_bytes(1000, 100) > 0


def test_argument_validation_without_gpu(lib_path):
    """EINVAL paths return before any CUDA call, so they are checkable on a CPU-only box."""
    from tuotuo_b200 import _native as nv

    lib = nv.load()
    rc = lib.ttlda_estep_f64(None, None, None, 1, 5, 10, 8, None, 1.0, None, None, None, None, None, None, None, 1,
                             1e-5, None, None, None, 0, None)
    assert rc == -1 and b"NULL" in lib.ttlda_last_error()
    rc = lib.ttlda_dirichlet_expectation_f64(None, 1, 1, 4, None, None, None)
    assert rc == -1


def test_no_cpu_fallback():
    """The product refuses to run without CUDA instead of silently computing on the host."""
    import numpy as np
    import torch

    from tuotuo_b200 import _native as nv
    from tuotuo_b200.lda_model import LDASmoothed

    with pytest.raises(nv.TTLDAError):
        nv.dirichlet_expectation(torch.ones(2, 4, dtype=torch.float64), 3, out_elog=torch.empty(2, 4, dtype=torch.float64))
    if not torch.cuda.is_available():
        with pytest.raises(nv.TTLDAError):
            LDASmoothed(3, verbose=False).fit(np.ones((4, 6)))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tuotuo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "lda_oracle" not in txt and "oracle/" not in txt and "import oracle" not in txt, f
