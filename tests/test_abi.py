"""CPU: the C-ABI library builds, loads and exports every symbol include/aec.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from async_ev_cnn_b200 import build, _native
    build.build_native()
    return _native.lib()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "aec.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aec_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    from async_ev_cnn_b200 import _native
    assert header_symbols() == sorted(_native.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in header_symbols():
        assert hasattr(lib, name), "libaec_b200.so does not export %s" % name
    assert lib.aec_version() >= 1000


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "aec.h")).read()
    assert "torch" not in text and "at::" not in text and "#include <cuda" not in text


def test_argument_errors_need_no_gpu(lib):
    from async_ev_cnn_b200 import _native
    h = ctypes.c_void_p()
    rc = lib.aec_net_create(ctypes.byref(h), 0, 0, 8, 8, 0.1, 0)
    assert rc == _native.AEC_EINVAL and b"n_streams" in lib.aec_last_error()
    assert lib.aec_net_step_host(None, None, None, 0, None, None) == _native.AEC_ESTATE


def test_product_path_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may import, load or exec it."""
    pkg = os.path.join(ROOT, "async-ev-cnn_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|import_module\([^)]*oracle|cutils_port|oracle[/.]\w+\.(py|so|c)\b", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not bad.search(src), "%s reaches into oracle/" % f
