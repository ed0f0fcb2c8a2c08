"""The C-ABI library loads and exports every symbol include/dsc.h declares (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from distropointclouds_b200 import build, _ffi
    build.build()
    return _ffi.load()


def _declared():
    src = open(os.path.join(ROOT, "include", "dsc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dsc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    from distropointclouds_b200 import _ffi
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dsc.h but not exported by libdsc.so"
    assert sorted(_ffi.EXPORTS) == names, "ctypes binding list out of sync with include/dsc.h"


def test_version_string(lib):
    assert lib.dsc_version().decode().startswith("dsc ")


def test_struct_layouts_match_header():
    from distropointclouds_b200 import _ffi
    assert C.sizeof(_ffi.DscParams) == 32
    assert C.sizeof(_ffi.DscScanStats) == 56
    assert _ffi.DscParams.max_length.offset == 8 and _ffi.DscParams.flags.offset == 28


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.dsc_create(0, C.byref(h))
    assert rc != 0 and not h.value
    from distropointclouds_b200.engine import Engine
    from distropointclouds_b200._ffi import DscError
    with pytest.raises(DscError):
        Engine(0)


def test_product_package_does_not_import_oracle():
    """The product path must never route through the CPU oracle."""
    pkg = os.path.join(ROOT, "distropointclouds_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "dsc_oracle" not in txt and "libdsc_oracle" not in txt, f
