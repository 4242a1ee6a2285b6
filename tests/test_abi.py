"""The C-ABI library loads and exports every symbol include/editanything_b200.h declares."""
import os
import re

from editanything_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "editanything_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ea_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by libea_b200.so"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)


def test_nothing_is_exported_that_the_header_does_not_declare():
    """The reverse direction: no undeclared `ea_*` entry point (debug hooks exist only in -DEA_GEMM_TIMING builds)."""
    import shutil
    import subprocess
    nm = shutil.which("nm")
    if nm is None:
        import pytest
        pytest.skip("binutils nm not available")
    out = subprocess.run([nm, "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.split() and ln.split()[-1].startswith("ea_")}
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "editanything_b200.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(ea_[a-z0-9_]+)\s*\(", hdr))
    assert exported <= declared, exported - declared


def test_library_reports_dtype_and_errors_without_gpu():
    lib = _lib.load()
    assert lib.ea_version() >= 1
    assert lib.ea_dtype_name() in (b"float16", b"bfloat16")
    assert b"argument" in lib.ea_strerror(-1)


def test_product_has_no_oracle_import():
    pkg = os.path.join(ROOT, "editanything_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
