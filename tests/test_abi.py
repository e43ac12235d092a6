"""CPU tests (-m "not gpu"): the C-ABI library loads, exports every symbol include/tnsb.h declares, and refuses to
compute without a GPU (no CPU fallback).  No compute calls are made here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tnsb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tnsb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_library):
    lib = C.CDLL(built_library)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tnsb.h but not exported by libtnsb.so"


def test_python_binding_covers_header(built_library):
    from treensearch_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    _lib.load()


def test_library_is_sm100a_and_has_no_oracle_dependency(built_library):
    out = subprocess.run(["cuobjdump", "-lelf", built_library], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    ldd = subprocess.run(["ldd", built_library], capture_output=True, text=True).stdout
    assert "tns_oracle" not in ldd and "tns_ref" not in ldd


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "treensearch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), f"{f} mentions the oracle"


def test_no_cpu_fallback(built_library):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import treensearch_b200 as t
    with pytest.raises(t.TreeNSearchError) as e:
        t.TreeNSearch()
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)
