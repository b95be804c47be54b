"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol
include/hyperion_b200.h declares, and refuses to run without a GPU (no silent CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from hyperion_b200 import capi
    return capi.load_library()


def _declared():
    text = open(os.path.join(ROOT, "include", "hyperion_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hyp_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 18
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_oracle_exports_checker_api():
    from oracle import oracle
    olib = oracle.load()
    for n in _declared():
        if n in ("hyp_version", "hyp_sizeof", "hyp_stream", "hyp_lucy_device_buffers", "hyp_lucy_photons", "hyp_ctx_create",
                 "hyp_finalize_setup", "hyp_run_lucy_iteration", "hyp_image_device_buffers"):
            continue
        assert hasattr(olib, "orc_" + n[4:]), n


def test_struct_layout_matches_header(lib):
    """The ctypes mirrors must have the C sizes (8-byte aligned structs of the header)."""
    from hyperion_b200 import capi
    for which, mirror in enumerate((capi.DustTables, capi.Source, capi.RunConf, capi.IterStats)):
        assert lib.hyp_sizeof(which) == C.sizeof(mirror), mirror


def test_no_gpu_is_a_loud_error(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hyperion_b200.capi import Engine, HyperionError
    with pytest.raises(HyperionError, match="no CUDA device|CPU fallback"):
        Engine(0)


def test_product_does_not_import_oracle():
    """Only tests/, smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "hyperion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.replace("the test oracle", "").replace("parity oracle", "") \
                    or f == "flatmodel.py", os.path.join(dirpath, f)
