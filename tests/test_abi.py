"""CPU-only: the CUDA C-ABI library loads and exports every symbol include/dfm_b200.h declares;
the product refuses to run without a device (no fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "dfm_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dfm_[a-z_0-9]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def libpath():
    from dynamic_factor_models_b200 import build
    return build.build()


def test_header_symbols_are_bound():
    from dynamic_factor_models_b200._lib import EXPORTS
    assert sorted(EXPORTS) == header_symbols()


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    for name in header_symbols():
        assert hasattr(lib, name), name
    lib.dfm_version.restype = ctypes.c_int
    assert lib.dfm_version() == 100
    lib.dfm_status_string.restype = ctypes.c_char_p
    assert lib.dfm_status_string(3) == b"matrix not positive definite"


def test_no_cpu_fallback(libpath):
    """Without a CUDA device handle creation must fail loudly (status 5), never fall back."""
    import torch
    from dynamic_factor_models_b200 import Library, DFMError
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(DFMError) as ei:
        Library(libpath)
    assert ei.value.code == 5


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "dynamic_factor_models_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src, fn


def test_shard_range(libpath):
    lib = ctypes.CDLL(libpath)
    lib.dfm_shard_range.argtypes = [ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong)]
    got = []
    for rank in range(8):
        b, e = ctypes.c_longlong(), ctypes.c_longlong()
        assert lib.dfm_shard_range(1000, rank, 8, ctypes.byref(b), ctypes.byref(e)) == 0
        got.append((b.value, e.value))
    assert got[0][0] == 0 and got[-1][1] == 1000
    assert all(got[i][1] == got[i + 1][0] for i in range(7))
    assert lib.dfm_shard_range(10, 8, 8, None, None) == 1


def test_header_is_plain_c99(tmp_path):
    """The boundary is a C ABI: include/dfm_b200.h must compile as C (what Julia's ccall / cgo / JNI stubs bind)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "t.c"
    src.write_text('#include "dfm_b200.h"\nint main(void) { dfm_em_opts o; dfm_factor_opts f; (void)o; (void)f; return DFM_OK; }\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
