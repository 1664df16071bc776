"""CPU: the C-ABI library loads and exports every symbol include/fastmot_b200.h declares."""
import os
import re
import ctypes

from conftest import ROOT


def _declared():
    hdr = open(os.path.join(ROOT, "include", "fastmot_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(fm_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported():
    names = _declared()
    assert len(names) >= 10
    so = os.path.join(ROOT, "fastmot_b200", "libfastmot_b200.so")
    assert os.path.exists(so), "run python -m fastmot_b200.build"
    lib = ctypes.CDLL(so)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_signatures_cover_header():
    from fastmot_b200 import _lib
    names = _declared()
    missing = [n for n in names if n not in _lib.SIGNATURES]
    assert not missing, missing
    _lib.load()
    assert _lib.load().fm_version() >= 100


def test_no_cpu_fallback_without_device():
    """Product classes must refuse to run without a B200 (no silent CPU path)."""
    import torch
    import pytest
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fastmot_b200 import _lib, MultiTracker
    with pytest.raises(_lib.FastMOTLibError):
        MultiTracker((1920, 1080), 'cosine')
