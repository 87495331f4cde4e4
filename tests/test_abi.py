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


def header_prototypes():
    """name -> list of C parameter types (pointer-ness and width) parsed from the header."""
    txt = (ROOT / "include" / "isoext_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    out = {}
    for ret, name, args in re.findall(r"\b([\w\s\*]+?)\b(isoext_\w+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        params = [a.strip() for a in args.replace("\n", " ").split(",")]
        out[name] = [] if params in ([""], ["void"]) else params
    return out


def _ctype_kind(t):
    """Coarse class of a ctypes argtype / a C parameter declaration: pointer, i64, i32, f32, size."""
    import ctypes as C
    if isinstance(t, str):
        if "*" in t:
            return "ptr"
        if "int64_t" in t or "uint64_t" in t:
            return "i64"
        if "size_t" in t:
            return "size"
        if "float" in t:
            return "f32"
        if re.search(r"\b(int|int32_t|uint32_t)\b", t):
            return "i32"
        raise AssertionError(f"unparsed parameter {t!r}")
    if t in (C.c_int64, C.c_uint64):
        return "i64"
    if t is C.c_size_t:
        return "size"
    if t is C.c_float:
        return "f32"
    if t in (C.c_int, C.c_int32, C.c_uint32):
        return "i32"
    return "ptr"      # c_void_p, c_char_p, POINTER(...)


def test_binding_argument_types_match_header():
    """Every entry of the ctypes table has the header's parameter count and, per position, the same class of type
    (pointer / 64-bit / 32-bit / float / size_t): an argument-order or width drift in the hand-kept table fails here
    (the header itself is compiler-checked against the definitions: every .cu includes it)."""
    from isoext_b200 import _lib
    protos = header_prototypes()
    assert sorted(protos) == sorted(_lib.SIGNATURES)
    for name, (res, argtypes) in _lib.SIGNATURES.items():
        want = [_ctype_kind(p) for p in protos[name]]
        got = [_ctype_kind(a) for a in argtypes]
        got = ["size" if (g == "i64" and w == "size") else g for g, w in zip(got, want)] + got[len(want):]
        assert got == want, f"{name}: ctypes {got} vs header {want}"


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
