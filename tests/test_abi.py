"""CPU tests of the drop-in boundary: the library loads and exports every symbol include/b200hist.h
declares (no compute calls -- there is no GPU here and no CPU fallback)."""
import ctypes
import os
import re

import pytest

from tests.conftest import ROOT


@pytest.fixture(scope="module")
def built_lib():
    from xgboost_b200 import _build
    return _build.build_library()


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "b200hist.h")).read()
    return sorted(set(re.findall(r"B200HIST_API\s+[\w\s\*]+?\b(b200hist_\w+)\s*\(", hdr)))


def test_header_declares_entry_points():
    syms = _declared_symbols()
    assert "b200hist_updater_update_tree" in syms and "b200hist_build_hist" in syms
    assert len(syms) >= 25


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for s in _declared_symbols():
        assert hasattr(lib, s), "missing export %s" % s


def test_python_binding_covers_the_header(built_lib):
    from xgboost_b200 import lib as L
    assert sorted(L.EXPORTED_SYMBOLS) == _declared_symbols()
    L.load()
    assert L.load().b200hist_version() == 100


def test_struct_layouts_match_the_header():
    from xgboost_b200 import lib as L
    assert ctypes.sizeof(L.TrainParam) == 64
    assert ctypes.sizeof(L.GPair64) == 16
    assert ctypes.sizeof(L.Quantiser) == 32
    assert ctypes.sizeof(L.Candidate) == 64
    assert ctypes.sizeof(L.TreeNode) == 48


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device every compute entry fails loudly instead of falling back."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import xgboost_b200 as xb
    cuts = xb.HistogramCuts([0, 2], [0.5, 1.5])
    with pytest.raises(xb.B200HistError):
        xb.EllpackPage(cuts, 10, 0, False)
    with pytest.raises(xb.B200HistError):
        xb.EllpackPage.from_dense(torch.zeros((4, 1)), cuts)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under xgboost_b200/ may reference it."""
    pkg = os.path.join(ROOT, "xgboost_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src \
                    or f == "_build.py", f


def test_only_fixture_generators_outside_tests_touch_the_oracle_directory():
    """Outside tests/, `oracle/` may appear in bench.py (cpu_baseline / --impl reference legs), in
    __graft_entry__.smoke() and in the FIXTURE GENERATORS under tools/ that drive the unmodified reference through
    oracle/ref_capi.py (the scripts that wrote tests/golden/*.json).  No tool may import the CPU restatement itself
    (`import oracle` / `oracle as orc`): checking against it is the tests' job."""
    allowed_ref_capi = {"make_golden_gpu.py", "make_golden_gpu_next.py", "make_golden_cpu_categorical.py",
                        "make_golden_agaricus.py", "probe_interaction_reference.py", "bench_predictor.py",
                        "bench_sparse.py"}
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tools")):
        for f in files:
            if not f.endswith(".py"):
                continue
            src = open(os.path.join(dirpath, f)).read()
            assert "import oracle" not in src and "liboracle" not in src, f
            if "from oracle" in src:
                assert f in allowed_ref_capi and "from oracle import ref_capi" in src, f
    entry = open(os.path.join(ROOT, "__graft_entry__.py")).read()
    build_body = entry[entry.index("def build()"):entry.index("def smoke()")]
    assert "import oracle" not in build_body and "from oracle" not in build_body
