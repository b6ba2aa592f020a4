"""The C-ABI library builds for sm_100a, loads, and exports exactly what include/tess_b200.h
declares; the ctypes binding covers the same set.  No compute calls (no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "tess_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tess_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from tess_b200 import build
    return build.build()


def test_header_declares_the_binding_surface():
    from tess_b200._capi import SIGNATURES
    assert declared_symbols() == sorted(SIGNATURES)


def test_library_exports_every_declared_symbol(lib_path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, missing


def test_library_loads_and_reports_version(lib_path):
    from tess_b200._capi import CApi
    api = CApi(lib_path)
    assert api.dll.tess_version() >= 100
    # no GPU in the CPU test environment: creating a context must fail cleanly, not crash
    import torch
    if not torch.cuda.is_available():
        h = ctypes.c_void_p()
        rc = api.dll.tess_ctx_create(0, ctypes.byref(h))
        assert rc < 0 and api.last_error()


def test_library_is_sm100a_only(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tess_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liblloyd_oracle" not in txt and "oracle/" not in txt, f


def test_engine_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from tess_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine()
    import numpy as np
    from tess_b200.lloyd import lloyd
    with pytest.raises(RuntimeError):
        lloyd(np.zeros((4, 2)), np.ones(4), np.zeros((2, 2)), 3)
