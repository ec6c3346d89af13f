"""The C-ABI library loads on a CPU-only box and exports every symbol include/ssr_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from simplesr_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ssr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ssr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    declared = set(_declared_symbols())
    bound = set(L.exported_symbols())
    assert declared == bound, (declared - bound, bound - declared)


def test_conv_desc_layout_matches_header():
    """ssr_conv_desc is 24 x 4-byte fields, passed by pointer; the mirror lists them in the header's order."""
    assert ctypes.sizeof(L.ConvDesc) == 24 * 4
    hdr = open(os.path.join(ROOT, "include", "ssr_b200.h")).read()
    body = hdr[hdr.index("typedef struct ssr_conv_desc {"):hdr.index("} ssr_conv_desc;")]
    import re
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [n.strip() for decl in re.findall(r"(?:int32_t|float)\s+([^;]+);", body) for n in decl.split(",")]
    assert names == [f[0] for f in L.ConvDesc._fields_], names


def test_no_cpu_fallback():
    """Without a B200 the context refuses to exist (the product path must fail loudly, never fall back)."""
    lib = L.load()
    h = ctypes.c_void_p()
    rc = lib.ssr_ctx_create(0, ctypes.byref(h))
    if rc == 0:
        lib.ssr_ctx_destroy(h)
        pytest.skip("a GPU is present")
    assert rc < 0
    assert b"no CUDA device" in lib.ssr_last_error() or b"sm_" in lib.ssr_last_error()
    with pytest.raises(Exception):
        L.Context(0)


def test_version_string():
    assert b"sm_100a" in L.load().ssr_version()


def test_product_package_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under simplesr_b200/ may import it."""
    pkg = os.path.join(ROOT, "simplesr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S), f
