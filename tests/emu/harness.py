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
    srcs = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(HERE, "cuda_emu.h"),
                                                              os.path.join(ROOT, "include", "tess_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-std=c++20", "-O2", "-pthread", "-DTESS_EMU", "-include", os.path.join(HERE, "cuda_emu.h"),
           "-x", "c++", os.path.join(SRC, "tess_b200.cu"), "-shared", "-fPIC", "-o", OUT]
    subprocess.check_call(cmd)
    return OUT


def _p(a):
    return None if a is None else a.ctypes.data


class EmuEngine(object):
    """NumPy-memory twin of tess_b200.engine.Engine (same call sequence, same C ABI)."""

    def __init__(self):
        import sys
        sys.path.insert(0, ROOT)
        from tess_b200._capi import CApi
        self.api = CApi(build())
        self.ctx = self.api.ctx_create(0)
        self.keep = {}
        self.n = 0
        self.k = 0

    def close(self):
        if self.ctx is not None:
            self.api.call("tess_ctx_destroy", self.ctx)
            self.ctx = None

    def set_option(self, opt, val):
        self.api.call("tess_set_option", self.ctx, opt, val)

    def set_image(self, density=None, signal=None, noise=None, x0=0.0, y0=0.0):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (density, signal, noise)]
        ref = next(a for a in arrs if a is not None)
        H, W = ref.shape
        self.keep["img"] = arrs
        self.n = H * W
        self.shape = (H, W)
        self.api.call("tess_set_image", self.ctx, H, W, float(x0), float(y0), _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), None)

    def set_points(self, xy, w):
        xy = np.ascontiguousarray(xy, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        self.keep["pts"] = (xy, w)
        self.n = xy.shape[0]
        self.shape = (self.n,)
        self.api.call("tess_set_points", self.ctx, self.n, _p(xy), _p(w), None)

    def set_generators(self, node, scale=None):
        node = np.ascontiguousarray(node, dtype=np.float64)
        sc = None if scale is None else np.ascontiguousarray(scale, dtype=np.float64)
        self.k = node.shape[0]
        self.api.call("tess_set_generators", self.ctx, self.k, _p(node), _p(sc), None)

    def step(self, n=1):
        self.api.call("tess_lloyd_step", self.ctx, n, None)

    def accumulate(self):
        self.api.call("tess_lloyd_accumulate", self.ctx, None)

    def update(self):
        self.api.call("tess_lloyd_update", self.ctx, None)

    def run(self, max_iters):
        return self.api.lloyd_run(self.ctx, max_iters, None)

    def labels(self):
        out = np.empty(self.n, dtype=np.int32)
        self.api.call("tess_get_labels", self.ctx, _p(out), None)
        return out.reshape(self.shape)

    def moments(self):
        out = np.empty((self.k, 6), dtype=np.float64)
        self.api.call("tess_get_moments", self.ctx, _p(out), None)
        return out

    def generators(self):
        out = np.empty((self.k, 2), dtype=np.float64)
        self.api.call("tess_get_generators", self.ctx, _p(out), None)
        return out

    def scales(self):
        out = np.empty(self.k, dtype=np.float64)
        self.api.call("tess_get_scales", self.ctx, _p(out), None)
        return out

    def stats(self):
        return self.api.iter_stats(self.ctx, None)

    def assign_grid(self, x0, y0, H, W, exhaustive=False):
        out = np.empty((H, W), dtype=np.int32)
        name = "tess_assign_grid_exhaustive" if exhaustive else "tess_assign_grid"
        self.api.call(name, self.ctx, float(x0), float(y0), H, W, _p(out), None)
        return out

    def assign_points(self, xy, exhaustive=False):
        xy = np.ascontiguousarray(xy, dtype=np.float64)
        out = np.empty(xy.shape[0], dtype=np.int32)
        name = "tess_assign_points_exhaustive" if exhaustive else "tess_assign_points"
        self.api.call(name, self.ctx, xy.shape[0], _p(xy), _p(out), None)
        return out
