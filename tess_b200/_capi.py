"""
tess_b200._capi -- ctypes binding of the C ABI in include/tess_b200.h.

This is the whole boundary between Python and the CUDA library: plain device pointers (ints),
sizes and a stream handle go in, status codes come out.  No torch types cross it.  The same
binding is what a tess maintainer would add on the reference side; the historic
precedent in the reference is examples/test_clloyds.py:21-39.
"""
import ctypes
import os

TESS_OK = 0
TESS_ERR_ARG = -1
TESS_ERR_CUDA = -2
TESS_ERR_STATE = -3
TESS_ERR_NONFINITE = -4
TESS_ERR_NOMEM = -5

MOM_COUNT, MOM_W, MOM_WX, MOM_WY, MOM_S, MOM_N2 = range(6)
OPT_WVT_SCALE_UPDATE = 1
OPT_UNIFORM_WEIGHT = 2
OPT_KERNEL_TIMING = 3
OPT_OWNER_UPDATE = 4

_vp = ctypes.c_void_p
_l = ctypes.c_long
_d = ctypes.c_double
_i = ctypes.c_int

# name -> (restype, argtypes); mirrors include/tess_b200.h one to one (tests/test_abi.py checks it)
SIGNATURES = {
    "tess_version": (_i, []),
    "tess_last_error": (ctypes.c_char_p, []),
    "tess_ctx_create": (_i, [_i, ctypes.POINTER(_vp)]),
    "tess_ctx_destroy": (_i, [_vp]),
    "tess_set_option": (_i, [_vp, _i, _l]),
    "tess_set_image": (_i, [_vp, _l, _l, _d, _d, _vp, _vp, _vp, _vp]),
    "tess_set_domain": (_i, [_vp, _d, _d, _d, _d]),
    "tess_set_points": (_i, [_vp, _l, _vp, _vp, _vp]),
    "tess_set_generators": (_i, [_vp, _l, _vp, _vp, _vp]),
    "tess_lloyd_accumulate": (_i, [_vp, _vp]),
    "tess_moments_ptr": (_i, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_l), ctypes.POINTER(_i)]),
    "tess_set_moment_buffer": (_i, [_vp, _vp, _l]),
    "tess_peer_allreduce": (_i, [_vp, _i, _i, ctypes.POINTER(_vp), _vp, _l, _vp]),
    "tess_peer_publish_range": (_i, [_vp, _vp]),
    "tess_peer_set_multicast": (_i, [_vp, _vp]),
    "tess_peer_allreduce_sparse": (_i, [_vp, _i, _i, ctypes.POINTER(_vp), _l, _vp]),
    "tess_lloyd_update": (_i, [_vp, _vp]),
    "tess_lloyd_step": (_i, [_vp, _l, _vp]),
    "tess_lloyd_run": (_i, [_vp, _l, ctypes.POINTER(_i), ctypes.POINTER(_l), _vp]),
    "tess_get_labels": (_i, [_vp, _vp, _vp]),
    "tess_get_moments": (_i, [_vp, _vp, _vp]),
    "tess_get_generators": (_i, [_vp, _vp, _vp]),
    "tess_get_scales": (_i, [_vp, _vp, _vp]),
    "tess_get_iter_stats": (_i, [_vp, ctypes.POINTER(_d), ctypes.POINTER(_l), ctypes.POINTER(_l), _vp]),
    "tess_get_status": (_i, [_vp, ctypes.POINTER(_i), ctypes.POINTER(_i), _vp]),
    "tess_get_kernel_timing": (_i, [_vp, ctypes.POINTER(_d), ctypes.POINTER(_l)]),
    "tess_launch_count": (_l, []),
    "tess_get_list_stats": (_i, [_vp, ctypes.POINTER(_l)]),
    "tess_assign_grid": (_i, [_vp, _d, _d, _l, _l, _vp, _vp]),
    "tess_assign_points": (_i, [_vp, _l, _vp, _vp, _vp]),
    "tess_label_sums": (_i, [_vp, _l, _vp, _vp, _l, _vp, _vp]),
    "tess_label_moments": (_i, [_vp, _l, _l, _vp, _vp, _l, _vp, _vp]),
    "tess_assign_grid_exhaustive": (_i, [_vp, _d, _d, _l, _l, _vp, _vp]),
    "tess_assign_points_exhaustive": (_i, [_vp, _l, _vp, _vp, _vp]),
}


class TessError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "libtess_b200 error %d: %s" % (code, message))
        self.code = code


class TessNonFinite(TessError, ValueError):
    """A generator became NaN/Inf -- the reference raises ValueError from cKDTree here
    (reference tess/lloyd.pyx:54 on the iteration after a zero-weight bin, lloyd.pyx:66-70)."""


class CApi(object):
    """The loaded library with typed entry points; every call raises on a non-zero status."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise OSError("libtess_b200 shared library not found at %s" % path)
        self.path = path
        self.dll = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.dll, name)      # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args

    def last_error(self):
        msg = self.dll.tess_last_error()
        return msg.decode("utf-8", "replace") if msg else ""

    def check(self, rc):
        if rc == TESS_OK:
            return
        msg = self.last_error()
        if rc == TESS_ERR_NONFINITE:
            raise TessNonFinite(rc, msg)
        if rc == TESS_ERR_ARG:
            raise ValueError("libtess_b200: " + msg)
        if rc == TESS_ERR_NOMEM:
            raise MemoryError("libtess_b200: " + msg)
        raise TessError(rc, msg)

    def call(self, name, *args):
        self.check(getattr(self.dll, name)(*args))

    # -- thin typed helpers ----------------------------------------------------------------
    def ctx_create(self, device):
        h = _vp()
        self.call("tess_ctx_create", int(device), ctypes.byref(h))
        return h

    def moments_ptr(self, ctx):
        p, n, m = _vp(), _l(), _i()
        self.call("tess_moments_ptr", ctx, ctypes.byref(p), ctypes.byref(n), ctypes.byref(m))
        return p.value, n.value, m.value

    def lloyd_run(self, ctx, max_iters, stream):
        conv, nit = _i(0), _l(0)
        self.call("tess_lloyd_run", ctx, int(max_iters), ctypes.byref(conv), ctypes.byref(nit), stream)
        return bool(conv.value), nit.value

    def iter_stats(self, ctx, stream):
        delta, nch, nit = _d(0), _l(0), _l(0)
        self.call("tess_get_iter_stats", ctx, ctypes.byref(delta), ctypes.byref(nch), ctypes.byref(nit), stream)
        return delta.value, nch.value, nit.value

    def status(self, ctx, stream):
        conv, bad = _i(0), _i(0)
        self.call("tess_get_status", ctx, ctypes.byref(conv), ctypes.byref(bad), stream)
        return bool(conv.value), bool(bad.value)

    def kernel_timing(self, ctx):
        ms, n = _d(0), _l(0)
        self.call("tess_get_kernel_timing", ctx, ctypes.byref(ms), ctypes.byref(n))
        return ms.value, n.value

    def launch_count(self):
        return int(self.dll.tess_launch_count())

    def list_stats(self, ctx):
        out = (_l * 4)()
        self.call("tess_get_list_stats", ctx, out)
        return dict(ring_fallback=out[0], arena_used=out[1], arena_capacity=out[2], scan_list=out[3])
