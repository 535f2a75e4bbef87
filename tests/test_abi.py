"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/eigb200.h declares; the ctypes prototypes cover exactly that set; no compute call is made."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "eigb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(eigb200_\w+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib_path():
    from eigensolver_gpu_b200 import build
    return build.build()


def test_header_declares_the_reference_entry_points():
    d = _declared()
    for name in ("eigb200_dsygvdx", "eigb200_zhegvdx", "eigb200_dsyevd", "eigb200_zheevd", "eigb200_init"):
        assert name in d


def test_library_exports_every_declared_symbol(lib_path):
    lib = C.CDLL(lib_path)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_prototypes_match_header():
    from eigensolver_gpu_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_strict_load_and_argument_counts(lib_path):
    from eigensolver_gpu_b200 import _lib
    _lib._lib = None
    lib = _lib.load(strict=True)
    assert lib.eigb200_version() >= 100
    txt = open(os.path.join(ROOT, "include", "eigb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    for name, (res, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", txt, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        assert n == len(args), (name, n, len(args))


def test_no_cpu_fallback_without_library(tmp_path, monkeypatch):
    from eigensolver_gpu_b200 import _lib
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(RuntimeError):
        _lib.load()


def test_product_does_not_import_oracle():
    import glob
    for p in glob.glob(os.path.join(ROOT, "eigensolver_gpu_b200", "**", "*.py"), recursive=True):
        src = open(p).read()
        assert "import oracle" not in src and "from oracle" not in src, p
