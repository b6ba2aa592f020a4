"""
tests/emu/harness.py -- TEST INFRASTRUCTURE.  Builds tess_b200/csrc/tess_b200.cu with g++ against
the SIMT emulator (tests/emu/cuda_emu.h) and drives the resulting library through the very same
ctypes binding as the product (tess_b200/_capi.py), with NumPy arrays standing in for device
memory.  It lets the kernel logic (candidate filters, warp choreography, convergence latch) be
checked against the oracle on the GPU-less build box.  The emulator is ~1e5x slower than a GPU:
keep inputs to a few tiles.  The product never loads this library.
"""
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "tess_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libtess_emu.so")


def build(force=False):
    # TESS_EMU_LIB: a pre-built emulator library (e.g. one compiled with -fsanitize=address,undefined by
    # tools/emu_sanitize.sh, which also preloads the sanitizer runtime)
    if os.environ.get("TESS_EMU_LIB"):
        return os.environ["TESS_EMU_LIB"]
    srcs = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(HERE, "cuda_emu.h"),
                                                              os.path.join(ROOT, "include", "tess_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-std=c++20", "-O2", "-pthread", "-DTESS_EMU", "-include", os.path.join(HERE, "cuda_emu.h"),
           "-x", "c++", os.path.join(SRC, "tess_b200.cu"), "-shared", "-fPIC", "-o", OUT]
    subprocess.check_call(cmd)
    return OUT


import ctypes
import sys

sys.path.insert(0, ROOT)
from tess_b200._capi import CApi          # noqa: E402
from tess_b200.engine import Engine       # noqa: E402

_API = None


def emu_api():
    global _API
    if _API is None:
        _API = CApi(build())
    return _API


class EmuEngine(Engine):
    """tess_b200.engine.Engine with NumPy memory and the emulator build of the library: the same
    host logic and the same C ABI calls as on the GPU, runnable on the GPU-less build box."""

    def __init__(self, device=None):
        self.api = emu_api()
        self.device = "emu"
        self.ctx = self.api.ctx_create(0)
        self._init_common()

    def _to_dev(self, a, dtype="float64"):
        return np.ascontiguousarray(a, dtype=dtype)

    def _empty(self, shape, dtype="float64"):
        return np.empty(shape, dtype=dtype)

    def _ptr(self, t):
        return None if t is None else t.ctypes.data

    def _stream(self):
        return None

    def _view(self, ptr, n):
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double)), shape=(n,))

    def _to_host(self, t):
        return t

    def _finite_mask(self, t):
        return np.isfinite(t)

    def _masked_fill(self, t, mask, value):
        t = t.copy()
        t[mask] = value
        return t

    def _amax(self, t):
        return int(t.max()) if t.size else -1
