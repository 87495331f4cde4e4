"""CPU: the C-ABI library loads and exports exactly the entry points include/isoext_b200.h declares,
and the Python binding table mirrors the header (no compute calls: there is no GPU here)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def header_functions():
    txt = (ROOT / "include" / "isoext_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(isoext_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from isoext_b200 import _lib
    handle = ctypes.CDLL(str(_lib.LIB_PATH))
    names = header_functions()
    assert len(names) >= 8
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"


def test_binding_table_matches_header():
    from isoext_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_functions()


def test_version_and_error_string_without_gpu():
    from isoext_b200 import _lib
    h = _lib.lib()
    assert h.isoext_abi_version() == 1
    assert b"sm_100a" in h.isoext_build_info()
    assert h.isoext_mc_dense_workspace_bytes(64, 64, 64, 1000) > 0
    # invalid shape: size query fails and leaves a message
    assert h.isoext_mc_dense_workspace_bytes(4, 4, 70000, 10) == 0
    assert b"65535" in h.isoext_last_error()


def test_product_never_imports_oracle():
    pkg = ROOT / "isoext_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        assert not re.search(r"\boracle\b", p.read_text()), f"{p} mentions oracle/"


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from isoext_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    import pytest
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.lib()
