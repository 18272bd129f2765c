"""The C-ABI library loads and exports every symbol include/aspire_b200.h
declares (no compute calls - there is no GPU here), and refuses to compute
without a device instead of falling back."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "aspire_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(asp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported():
    from aspire_b200 import _lib

    names = _declared()
    assert len(names) >= 30
    assert sorted(_lib.SIGNATURES) == names  # the ctypes table mirrors the header
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), n
    assert lib.asp_version() == 100


def test_no_cpu_fallback():
    import torch

    from aspire_b200 import AspireB200Error, _lib, ops

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert _lib.load().asp_device_check() == _lib.ASP_ERR_DEVICE
    with pytest.raises(AspireB200Error):
        ops.device_check()
    with pytest.raises(AspireB200Error):
        ops.reweight(torch.zeros(4, dtype=torch.float64), torch.zeros(4, dtype=torch.float64),
                     torch.zeros(4, dtype=torch.float64))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "aspire_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
