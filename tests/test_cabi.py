"""CPU: the C-ABI shared library builds, loads and exports every symbol include/b200nerf.h declares.
No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200nerf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200nerf_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from neurad_studio_b200 import lib

    assert set(_declared_symbols()) == set(lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    from neurad_studio_b200 import build, lib

    path = build.build()
    cdll = ctypes.CDLL(path)
    for name in _declared_symbols():
        assert hasattr(cdll, name), f"{name} not exported"
    bound = lib.load(build_if_missing=False)
    assert bound.b200nerf_version() == 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from neurad_studio_b200.backend import B200Backend

    with pytest.raises(RuntimeError):
        B200Backend()
    # and the C library itself reports the failure instead of computing on the host
    from neurad_studio_b200 import lib

    l = lib.load()
    h = ctypes.c_void_p()
    assert l.b200nerf_create(0, ctypes.byref(h)) != 0
    assert l.b200nerf_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "neurad-studio_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dirpath, f)
