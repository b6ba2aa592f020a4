"""
tess_b200.engine -- one Lloyd / Voronoi-assignment context on one GPU.

`Engine` owns a `tess_ctx` of libtess_b200.so (include/tess_b200.h) and keeps the torch CUDA
tensors whose device pointers it lent to the library alive.  torch is used for device memory and
streams only; every computation happens in the library's CUDA kernels.  If the library is missing
or no CUDA device is present, construction raises -- there is no CPU path.

The memory hooks (`_to_dev`, `_empty`, `_ptr`, `_stream`, `_view`, `_to_host`) are the only places
that touch torch, so that the CPU test-suite can drive the identical host logic through the SIMT
emulator build of the same C ABI (tests/emu/harness.py) with NumPy memory.
"""
import numpy as np

from . import _capi


class _RawCudaArray(object):
    """__cuda_array_interface__ carrier for a raw device pointer (used to alias the library's
    moment vector as a torch tensor without a copy)."""

    def __init__(self, ptr, n, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class Engine(object):
    """Lloyd / WVT iteration state for one pixel map or point set on one device.

    Call order: set_image(...) or set_points(...); set_generators(...); then step()/run() and the
    getters.  Mirrors the C ABI one to one (include/tess_b200.h)."""

    def __init__(self, device=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("tess_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        from ._lib import get_api
        self._torch = torch
        self.api = get_api()
        if device is None:
            device = torch.cuda.current_device()
        if not isinstance(device, int):
            device = torch.device(device).index          # 'cuda' / index-less: the current device, not GPU 0
            if device is None:
                device = torch.cuda.current_device()
        self.device = torch.device("cuda", device)
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.empty(1, device=self.device)          # make sure the primary context exists
            self.ctx = self.api.ctx_create(self.device.index)
        self._init_common()

    def _init_common(self):
        self._keep = {}
        # whose generator table the context holds (several tessellation objects may share one Engine):
        # reset by every call that rewrites the table, claimed by the object that uploaded it
        self.gen_owner = None
        self.shape = None
        self.n = 0
        self.k = 0
        self.source = None

    # ------------------------------------------------------------------ memory hooks
    def _to_dev(self, a, dtype="float64"):
        torch = self._torch
        tdt = getattr(torch, dtype)
        if isinstance(a, torch.Tensor):
            t = a.to(device=self.device, dtype=tdt, non_blocking=True)
        else:
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(self.device, non_blocking=True)
        return t.contiguous()

    def _empty(self, shape, dtype="float64"):
        return self._torch.empty(shape, dtype=getattr(self._torch, dtype), device=self.device)

    def _ptr(self, t):
        return None if t is None else t.data_ptr()

    def _stream(self):
        return self._torch.cuda.current_stream(self.device).cuda_stream

    def _view(self, ptr, n):
        return self._torch.as_tensor(_RawCudaArray(ptr, n), device=self.device)

    def _to_host(self, t):
        return t.cpu().numpy()

    def _finite_mask(self, t):
        return self._torch.isfinite(t)

    def _masked_fill(self, t, mask, value):
        return t.masked_fill(mask, value)

    def _amax(self, t):
        return int(t.max().item()) if t.numel() else -1

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "ctx", None) is not None:
            self.api.call("tess_ctx_destroy", self.ctx)
            self.ctx = None
            self._keep = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ inputs
    def set_option(self, option, value):
        self.api.call("tess_set_option", self.ctx, int(option), int(value))

    def set_image(self, density=None, signal=None, noise=None, x0=0.0, y0=0.0):
        """Pixel map mode (replaces the point-set construction of CVTessellation.from_image,
        reference tess/voronoi.py:281-287).  Either `density` [H,W] or `signal` and `noise`."""
        arrs = [None if a is None else self._to_dev(a) for a in (density, signal, noise)]
        ref = next((a for a in arrs if a is not None), None)
        if ref is None or ref.ndim != 2:
            raise ValueError("set_image: need a 2-D density map or 2-D signal and noise maps")
        for a in arrs:
            if a is not None and tuple(a.shape) != tuple(ref.shape):
                raise ValueError("set_image: maps must have the same shape")
        H, W = int(ref.shape[0]), int(ref.shape[1])
        self.api.call("tess_set_image", self.ctx, H, W, float(x0), float(y0),
                      self._ptr(arrs[0]), self._ptr(arrs[1]), self._ptr(arrs[2]), self._stream())
        self._keep["input"] = arrs
        self.shape, self.n, self.source = (H, W), H * W, "image"

    def set_domain(self, x0, y0, x1, y1):
        """Rectangle containing (nearly) all generators: the generator grid covers it (a row-sharded
        rank passes the full image; see include/tess_b200.h)."""
        self.api.call("tess_set_domain", self.ctx, float(x0), float(y0), float(x1), float(y1))

    def set_points(self, xy, w):
        """Point-set mode: the `xy`, `w` arguments of lloyd() (reference tess/lloyd.pyx:10-11)."""
        xy = self._to_dev(xy)
        w = self._to_dev(w)
        if xy.ndim != 2 or xy.shape[1] != 2 or w.ndim != 1 or w.shape[0] != xy.shape[0]:
            raise ValueError("set_points: xy must be (n, 2) and w (n,)")
        n = int(xy.shape[0])
        self.api.call("tess_set_points", self.ctx, n, self._ptr(xy), self._ptr(w), self._stream())
        self._keep["input"] = [xy, w]
        self.shape, self.n, self.source = (n,), n, "points"

    def set_generators(self, node_xy, scale=None):
        node = self._to_dev(node_xy)
        if node.ndim != 2 or node.shape[1] != 2:
            raise ValueError("set_generators: node_xy must be (k, 2)")
        sc = None if scale is None else self._to_dev(scale)
        if sc is not None and (sc.ndim != 1 or sc.shape[0] != node.shape[0]):
            raise ValueError("set_generators: scale must be (k,)")
        self.k = int(node.shape[0])
        self.gen_owner = None
        self.api.call("tess_set_generators", self.ctx, self.k, self._ptr(node), self._ptr(sc), self._stream())

    # ------------------------------------------------------------------ iteration
    def accumulate(self):
        self.api.call("tess_lloyd_accumulate", self.ctx, self._stream())

    def update(self):
        self.gen_owner = None
        self.api.call("tess_lloyd_update", self.ctx, self._stream())

    def step(self, n=1):
        self.gen_owner = None
        self.api.call("tess_lloyd_step", self.ctx, int(n), self._stream())

    def run(self, max_iters):
        """Reference loop and stop rule (tess/lloyd.pyx:47-90).  Returns (converged, n_iters)."""
        self.gen_owner = None
        return self.api.lloyd_run(self.ctx, int(max_iters), self._stream())

    def set_moment_buffer(self, t, n_doubles):
        """Keep the packed moment vector in caller memory `t` (a symmetric allocation shared with the
        peer ranks, see tess_b200.sharded); None returns to a library-owned vector."""
        self.api.call("tess_set_moment_buffer", self.ctx, None if t is None else self._ptr(t), int(n_doubles))
        self._keep["mom_ext"] = t

    def peer_allreduce(self, world, rank, peer_ptrs, multicast_ptr, n_doubles):
        """Sum the moment vector over the ranks in place with the library's own kernel
        (tess_peer_allreduce): `peer_ptrs` a ctypes array of the peer-mapped addresses."""
        self.api.call("tess_peer_allreduce", self.ctx, int(world), int(rank), peer_ptrs,
                      multicast_ptr if multicast_ptr else None, int(n_doubles), self._stream())

    def peer_publish_range(self):
        """First / last bin with a non-zero local count -> the two slots behind the caller's moment
        buffer (tess_peer_publish_range); read by the peers' sparse exchange."""
        self.api.call("tess_peer_publish_range", self.ctx, self._stream())

    def peer_set_multicast(self, multicast_ptr):
        """tess_peer_set_multicast: the sparse exchange broadcasts through the NVLS multicast mapping
        (0 / None: peer-to-peer copies)."""
        self.api.call("tess_peer_set_multicast", self.ctx, multicast_ptr if multicast_ptr else None)

    def peer_allreduce_sparse(self, world, rank, peer_ptrs, n_doubles):
        """tess_peer_allreduce_sparse: rows are read only from the peers whose published range holds them."""
        self.api.call("tess_peer_allreduce_sparse", self.ctx, int(world), int(rank), peer_ptrs, int(n_doubles),
                      self._stream())

    def moments_view(self):
        """The library's packed local moment vector [k*M | tail] as a tensor ALIASING device memory:
        what a multi-GPU driver all-reduces between accumulate() and update()."""
        ptr, n, m = self.api.moments_ptr(self.ctx)
        return self._view(ptr, n), m

    # ------------------------------------------------------------------ results
    def labels(self):
        out = self._empty(self.shape, "int32")
        self.api.call("tess_get_labels", self.ctx, self._ptr(out), self._stream())
        return out

    def moments(self):
        out = self._empty((self.k, 6))
        self.api.call("tess_get_moments", self.ctx, self._ptr(out), self._stream())
        return out

    def generators(self):
        out = self._empty((self.k, 2))
        self.api.call("tess_get_generators", self.ctx, self._ptr(out), self._stream())
        return out

    def scales(self):
        out = self._empty((self.k,))
        self.api.call("tess_get_scales", self.ctx, self._ptr(out), self._stream())
        return out

    def stats(self):
        """(delta of the last update, labels changed in the last accumulate or -1, iterations done)."""
        return self.api.iter_stats(self.ctx, self._stream())

    def kernel_timing(self):
        """(summed device ms, launches) of the dominant kernel since the last call (bench.py)."""
        return self.api.kernel_timing(self.ctx)

    def list_stats(self):
        """Diagnostics of the last candidate-list build (see tess_get_list_stats)."""
        return self.api.list_stats(self.ctx)

    def status(self):
        """(converged latch, a generator became non-finite)."""
        return self.api.status(self.ctx, self._stream())

    # ------------------------------------------------------------------ assignment only
    def assign_grid(self, x0, y0, H, W, exhaustive=False):
        out = self._empty((int(H), int(W)), "int32")
        name = "tess_assign_grid_exhaustive" if exhaustive else "tess_assign_grid"
        self.api.call(name, self.ctx, float(x0), float(y0), int(H), int(W), self._ptr(out), self._stream())
        return out

    def assign_points(self, xy, exhaustive=False):
        xy = self._to_dev(xy)
        if xy.ndim != 2 or xy.shape[1] != 2:
            raise ValueError("assign_points: xy must be (m, 2)")
        out = self._empty((int(xy.shape[0]),), "int32")
        name = "tess_assign_points_exhaustive" if exhaustive else "tess_assign_points"
        self.api.call(name, self.ctx, int(xy.shape[0]), self._ptr(xy), self._ptr(out), self._stream())
        self._keep["query"] = xy
        return out

    def label_moments(self, labels, weights=None, k=None):
        """(count, sum w, sum w*row, sum w*col) per label of a 2-D int32 label image, pixel coordinates
        implicit, one pass (tess_label_moments; PixelAccretor.centroids, pixel_accretion.py:212-240)."""
        lab = self._to_dev(labels, "int32")
        if lab.ndim != 2:
            raise ValueError("label_moments: labels must be a 2-D image")
        w = None if weights is None else self._to_dev(weights)
        if w is not None and tuple(w.shape) != tuple(lab.shape):
            raise ValueError("label_moments: weights and labels differ in shape")
        k = self.k if k is None else int(k)
        out = self._empty((k, 4))
        self.api.call("tess_label_moments", self.ctx, int(lab.shape[0]), int(lab.shape[1]), self._ptr(lab), self._ptr(w), k,
                      self._ptr(out), self._stream())
        self._keep["sums"] = (lab, w)
        return out

    def label_sums(self, labels, weights=None, k=None):
        """Per-bin sums of `weights` (or counts) over an int32 label array -- np.bincount of the
        reference's compute_cell_areas / sum_cell_point_mass (tess/voronoi.py:150,195)."""
        lab = self._to_dev(labels, "int32").reshape(-1)
        w = None if weights is None else self._to_dev(weights).reshape(-1)
        if w is not None and w.shape[0] != lab.shape[0]:
            raise ValueError("label_sums: weights and labels differ in length")
        k = self.k if k is None else int(k)
        out = self._empty((k,))
        self.api.call("tess_label_sums", self.ctx, int(lab.shape[0]), self._ptr(lab), self._ptr(w), k,
                      self._ptr(out), self._stream())
        self._keep["sums"] = (lab, w)
        return out
