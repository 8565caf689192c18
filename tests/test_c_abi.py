"""The drop-in boundary: libhammlet_b200.so must load and export every function that include/hammlet_b200.h and
include/hammlet_host.h declare, the ctypes binding must name every one of them, and — there is no CPU fallback —
creating a handle without a CUDA device must fail with a message instead of computing anything on the host."""
import ctypes as C
import os
import re

import pytest

from hammlet_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)          # comments
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)            # preprocessor lines
    return sorted(set(re.findall(r"\b((?:hml|hammlet)_[a-z0-9_]+)\s*\(", text)))


@pytest.mark.parametrize("header", ["hammlet_b200.h", "hammlet_host.h"])
def test_every_declared_function_is_exported_and_bound(header):
    names = declared(header)
    assert len(names) >= 8
    lib = C.CDLL(capi.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/{header} but not exported by libhammlet_b200.so"
    missing = [n for n in names if n not in capi.EXPORTS]
    assert not missing, f"capi.EXPORTS (checked by load_library) lacks {missing}"


def test_binding_names_only_declared_functions():
    both = set(declared("hammlet_b200.h")) | set(declared("hammlet_host.h"))
    assert set(capi.EXPORTS) <= both


def test_load_library_resolves_everything():
    lib = capi.load_library()
    assert lib.hml_version().decode().startswith("hammlet_b200")


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.HmlError) as e:
        capi.Handle(0)
    assert "CUDA" in str(e.value) or "device" in str(e.value)
