"""
oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Three CPU views of the same algorithm, used only as checkers / CPU baseline:

  1. `c_*`      ctypes bindings of the plain-C restatement oracle/lloyd_oracle.c
  2. `np_*`     NumPy restatement (brute-force argmin with the pinned d^2 formula, np.bincount
                sums) -- SURVEY.md section 8c: bit-identical to the compiled reference
  3. `ref_*`    the compiled reference itself (oracle/_ref/lloyd*.so = tess/lloyd.pyx built
                unmodified) when present; plus `kd_assign` = scipy cKDTree, the third-party
                code the reference calls at tess/lloyd.pyx:54-55.

Pinned arithmetic (all fp64, no FMA):
  d2(i,j)  = fl( fl(dx*dx) + fl(dy*dy) ),  dx = x_i - gx_j, dy = y_i - gy_j   [lloyd.pyx:55 via cKDTree]
  label_i  = argmin_j d2(i,j)            ties -> lowest j (restatement's rule; cKDTree's is
                                         tree-order dependent, so fixtures are tie-free)
  WVT ext. = argmin_j fl( d2(i,j) * inv_s2_j ),  inv_s2_j = fl(1 / fl(s_j*s_j))   [no reference counterpart]
  moments  : count, sum w, sum fl(w*x), sum fl(w*y) sequential in point order [lloyd.pyx:58-65]
             (+ sum S, sum fl(N*N): per-bin S/N = sum S / sqrt(sum N^2), pixel_accretion.py:501-517)
  node     = (sum wx / sum w, sum wy / sum w) if count > 0 else (0, 0)      [lloyd.pyx:66-74]
  delta    = sum_j fl(fl(dx*dx)+fl(dy*dy)) sequential over all j           [lloyd.pyx:76-81]
  stop     : delta == 0 -> True; n_iters > max_iters -> False                [lloyd.pyx:85-90]
"""
import ctypes
import glob
import io
import os
import subprocess
import sys
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


# ----------------------------------------------------------------------------- C restatement
def build_c(force=False):
    so = os.path.join(HERE, "liblloyd_oracle.so")
    src = os.path.join(HERE, "lloyd_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "-B", "liblloyd_oracle.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c())
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int64)
        lp = ctypes.POINTER(ctypes.c_long)
        lib.oracle_assign_points.argtypes = [ctypes.c_long, dp, ctypes.c_long, dp, dp, ip]
        lib.oracle_assign_points.restype = None
        lib.oracle_assign_grid.argtypes = [ctypes.c_long, ctypes.c_long, ctypes.c_double,
                                           ctypes.c_double, ctypes.c_long, dp, dp, ip]
        lib.oracle_assign_grid.restype = None
        lib.oracle_accumulate.argtypes = [ctypes.c_long, dp, dp, ip, dp, dp, ctypes.c_long, dp]
        lib.oracle_accumulate.restype = None
        lib.oracle_update_nodes.argtypes = [ctypes.c_long, dp, dp, dp]
        lib.oracle_update_nodes.restype = ctypes.c_double
        lib.oracle_lloyd.argtypes = [ctypes.c_long, dp, dp, ctypes.c_long, dp, ctypes.c_long,
                                     ip, lp, dp, dp]
        lib.oracle_lloyd.restype = ctypes.c_int
        lib.oracle_num_threads.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def _dp(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def inv_scale2(scale):
    """inv_s2_j = fl(1 / fl(s_j*s_j)) -- the WVT extension's pinned form."""
    if scale is None:
        return None
    s = _c64(scale)
    return 1.0 / (s * s)


def c_assign_points(xy, node, scale=None):
    xy = _c64(xy); node = _c64(node); inv = inv_scale2(scale)
    idx = np.empty(xy.shape[0], dtype=np.int64)
    _lib().oracle_assign_points(xy.shape[0], _dp(xy), node.shape[0], _dp(node), _dp(inv),
                                idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    return idx


def c_assign_grid(H, W, node, x0=0.0, y0=0.0, scale=None):
    node = _c64(node); inv = inv_scale2(scale)
    lab = np.empty((H, W), dtype=np.int64)
    _lib().oracle_assign_grid(H, W, float(x0), float(y0), node.shape[0], _dp(node), _dp(inv),
                              lab.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    return lab


def c_accumulate(xy, w, idx, k, signal=None, noise=None):
    xy = _c64(xy); w = _c64(w); idx = np.ascontiguousarray(idx, dtype=np.int64)
    s = None if signal is None else _c64(signal)
    nz = None if noise is None else _c64(noise)
    mom = np.zeros((k, 6), dtype=np.float64)
    _lib().oracle_accumulate(xy.shape[0], _dp(xy), _dp(w),
                             idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                             _dp(s), _dp(nz), k, _dp(mom))
    return mom


def c_update_nodes(mom, old_node):
    mom = _c64(mom); old = _c64(old_node)
    new = np.empty_like(old)
    delta = _lib().oracle_update_nodes(old.shape[0], _dp(mom), _dp(old), _dp(new))
    return new, delta


def c_lloyd(xy, w, node_xy, max_iters):
    """Full loop, C restatement.  Returns (node_xy, idx, converged, n_iters, deltas, moments)."""
    xy = _c64(xy); w = _c64(w); node = _c64(node_xy).copy()
    n, k = xy.shape[0], node.shape[0]
    idx = np.empty(n, dtype=np.int64)
    deltas = np.zeros(max(int(max_iters), 0) + 2, dtype=np.float64)
    mom = np.zeros((k, 6), dtype=np.float64)
    nit = ctypes.c_long(0)
    conv = _lib().oracle_lloyd(n, _dp(xy), _dp(w), k, _dp(node), int(max_iters),
                               idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                               ctypes.byref(nit), _dp(deltas), _dp(mom))
    return node, idx, bool(conv), int(nit.value), deltas[:nit.value], mom


# ----------------------------------------------------------------------------- NumPy restatement
def np_assign_points(xy, node, scale=None, chunk=4096):
    """Brute-force argmin with the pinned formula; np.argmin returns the lowest index on ties."""
    xy = _c64(xy); node = _c64(node); inv = inv_scale2(scale)
    out = np.empty(xy.shape[0], dtype=np.int64)
    gx, gy = node[:, 0][None, :], node[:, 1][None, :]
    for s in range(0, xy.shape[0], chunk):
        dx = xy[s:s + chunk, 0][:, None] - gx
        dy = xy[s:s + chunk, 1][:, None] - gy
        d = dx * dx + dy * dy          # separately rounded products, then one add
        if inv is not None:
            d = d * inv[None, :]
        out[s:s + chunk] = np.argmin(d, axis=1)
    return out


def np_moments(xy, w, idx, k, signal=None, noise=None):
    """np.bincount sums sequentially in index order -> bitwise equal to lloyd.pyx:58-65."""
    xy = _c64(xy); w = _c64(w)
    good = idx >= 0
    i = idx[good]
    mom = np.zeros((k, 6), dtype=np.float64)
    mom[:, 0] = np.bincount(i, minlength=k)
    mom[:, 1] = np.bincount(i, weights=w[good], minlength=k)
    mom[:, 2] = np.bincount(i, weights=w[good] * xy[good, 0], minlength=k)
    mom[:, 3] = np.bincount(i, weights=w[good] * xy[good, 1], minlength=k)
    if signal is not None:
        mom[:, 4] = np.bincount(i, weights=_c64(signal)[good], minlength=k)
    if noise is not None:
        nz = _c64(noise)[good]
        mom[:, 5] = np.bincount(i, weights=nz * nz, minlength=k)
    return mom


def np_update_nodes(mom, old_node):
    new = np.zeros_like(old_node, dtype=np.float64)
    occ = mom[:, 0] > 0
    with np.errstate(invalid="ignore", divide="ignore"):
        new[occ, 0] = mom[occ, 2] / mom[occ, 1]
        new[occ, 1] = mom[occ, 3] / mom[occ, 1]
    d = (old_node[:, 0] - new[:, 0]) ** 2 + (old_node[:, 1] - new[:, 1]) ** 2
    delta = 0.0
    for v in d:                      # sequential, like lloyd.pyx:78-81
        delta += v
    return new, float(delta)


def np_lloyd(xy, w, node_xy, max_iters, assign=None, record=None):
    """Full loop (lloyd.pyx:47-90).  `assign` defaults to the brute-force restatement;
    `record` (a list) receives per-iteration dicts(labels, moments, nodes_in, nodes_out, delta)."""
    assign = assign or np_assign_points
    node = _c64(node_xy).copy()
    n_iters = 0
    while True:
        idx = assign(xy, node)
        mom = np_moments(xy, w, idx, node.shape[0])
        new, delta = np_update_nodes(mom, node)
        if record is not None:
            record.append(dict(labels=idx.copy(), moments=mom, nodes_in=node.copy(),
                               nodes_out=new.copy(), delta=delta))
        node = new
        if delta == 0:
            return node, idx, True
        elif n_iters > max_iters:
            return node, idx, False
        n_iters += 1


def image_points(density):
    """Point set of tess/voronoi.py:281-287: row-major pixel centres (x=col, y=row), finite only.
    Returns (xy[n_good,2], w[n_good], good_flat_index)."""
    H, W = density.shape
    x, y = np.meshgrid(np.arange(W, dtype=float), np.arange(H, dtype=float))
    xy = np.column_stack((x.ravel(), y.ravel()))
    dens = np.asarray(density, dtype=np.float64).ravel()
    good = np.where(np.isfinite(dens))[0]
    return xy[good], dens[good], good


def bin_sn(mom):
    """Per-bin S/N = sum S / sqrt(sum N^2)  (tess/pixel_accretion.py:513-514)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        return mom[:, 4] / np.sqrt(mom[:, 5])


# ----------------------------------------------------------------------------- the reference itself
def kd_assign(xy, node, scale=None):
    """scipy cKDTree 1-NN: the third-party call the reference makes at tess/lloyd.pyx:54-55."""
    assert scale is None
    from scipy.spatial import cKDTree
    return cKDTree(node).query(xy, k=1)[1].astype(np.int64)


def ref_available():
    return bool(glob.glob(os.path.join(HERE, "_ref", "lloyd.*.so")))


def ref_module():
    """Import the compiled reference tess/lloyd.pyx from oracle/_ref (built by oracle/build_ref.py)."""
    import scipy.spatial  # noqa: F401  (the .pyx imports it at module import)
    p = os.path.join(HERE, "_ref")
    if p not in sys.path:
        sys.path.insert(0, p)
    import lloyd as _ref_lloyd
    return _ref_lloyd


def ref_lloyd(xy, w, node_xy, max_iters, want_stdout=False):
    """Run the compiled reference lloyd() (unmodified tess/lloyd.pyx:10-90).
    Its per-iteration `print` (lloyd.pyx:82) is captured and returned when requested."""
    mod = ref_module()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        node, idx, conv = mod.lloyd(_c64(xy), _c64(w), _c64(node_xy), int(max_iters))
    out = (np.asarray(node).copy(), np.asarray(idx).astype(np.int64), bool(conv))
    return out + (buf.getvalue(),) if want_stdout else out
