"""CPU: the C-ABI library loads and exports every symbol include/fnx.h declares (no compute calls)."""
import os
import re

from conftest import ROOT


def _declared():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            txt = open(os.path.join(ROOT, "include", fn)).read()
            txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
            names |= set(re.findall(r"\b(fnx_[a-z0-9_]+)\s*\(", txt))
    return names


def test_header_symbols_exported(libfnx):
    from fluidnexus_b200 import _lib
    declared = _declared()
    assert len(declared) >= 10
    for name in sorted(declared):
        assert hasattr(libfnx, name), f"libfnx.so does not export {name}"
    # and the ctypes table binds exactly the declared functions
    assert set(_lib.SYMBOLS) == declared, set(_lib.SYMBOLS) ^ declared


def test_versions(libfnx):
    assert libfnx.fnx_abi_version() == 1
    assert libfnx.fnx_build_arch() == b"sm_100a"


def test_scratch_sizes_monotone(libfnx):
    g1, g2 = libfnx.fnx_raster_geom_bytes(1000, 1), libfnx.fnx_raster_geom_bytes(2000, 5)
    assert 0 < g1 < g2
    b1, b2 = libfnx.fnx_raster_binning_bytes(10_000, 1), libfnx.fnx_raster_binning_bytes(10_000, 3)
    assert 0 < b1 < b2
    assert libfnx.fnx_raster_image_bytes(64, 64, 1) < libfnx.fnx_raster_image_bytes(512, 512, 5)


def test_only_sm100a_code_in_library():
    import subprocess
    from fluidnexus_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import importlib
    from fluidnexus_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    import pytest
    with pytest.raises(ImportError):
        _lib.lib()
