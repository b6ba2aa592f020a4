"""
tess_b200.voronoi -- drop-in for the reference's `tess.voronoi` (tess/voronoi.py:31-321):
`VoronoiTessellation` and `CVTessellation` with the reference's attribute and method names.

All nearest-generator queries (segmap, render_voronoi_field, partition_points -- the reference's
`griddata(..., method='nearest')` at voronoi.py:113 and `KDTree.query` at voronoi.py:171-172) and all
per-cell sums (np.bincount at voronoi.py:150,195, node_weights at voronoi.py:312-321) run in
libtess_b200.so on the GPU.  `CVTessellation.from_image` never materialises the (n, 2) pixel
coordinate array the reference builds at voronoi.py:281-283.

Array types follow the inputs: NumPy in -> NumPy out, torch CUDA in -> torch CUDA out.
Label arrays are int64 at this surface (the reference's `long`); `segmap` is int64 as in the
SciPy of the reference's time (modern SciPy returns float64 from griddata).
"""
import logging
import math

import numpy as np

from .engine import Engine
from .lloyd import _check_f64, _is_torch, run_lloyd

log = logging.getLogger(__name__)


def _np(a):
    return a.detach().cpu().numpy() if _is_torch(a) else np.asarray(a)


class VoronoiTessellation(object):
    """Voronoi tessellation of the plane given by its nodes (generators); GPU-backed.

    Parameters
    ----------
    xy : (n_nodes, 2) float64 array, NumPy or torch CUDA
        ``(x, y)`` of every node.
    """
    def __init__(self, xy, device=None, engine=None):
        super(VoronoiTessellation, self).__init__()
        self._xy = xy
        self._segmap = None      #: 2D array of bin numbers for each pixel
        self._cell_areas = None  #: 1D array of Voronoi cell areas
        self.xlim = None         #: ``(min, max)`` coords of x pixel grid
        self.ylim = None         #: ``(min, max)`` coords of y pixel grid
        self._device = device
        self._engine = engine
        self._torch_out = _is_torch(xy)

    # -- engine plumbing -------------------------------------------------------------------
    def _eng_raw(self):
        """The engine, without making sure that it holds this object's generators."""
        if self._engine is None:
            dev = self._device
            if dev is None and _is_torch(self._xy) and self._xy.is_cuda:
                dev = self._xy.device
            self._engine = Engine(dev)
        return self._engine

    def _eng(self):
        self._eng_raw()
        # several objects may share one Engine (the public `engine=` argument): upload this
        # object's table unless it is the one the context holds
        if self._engine.gen_owner is not self:
            self._engine.set_generators(self._xy)
            self._engine.gen_owner = self
        return self._engine

    def _out(self, t, dtype=None):
        """Device tensor -> the caller's array family."""
        if self._torch_out:
            return t if dtype is None else t.to(dtype=getattr(self._eng_raw()._torch, dtype))
        a = self._eng_raw()._to_host(t)
        return a if dtype is None else a.astype(dtype)

    # -- reference surface -----------------------------------------------------------------
    @property
    def nodes(self):
        """The ``(n_nodes, 2)`` node coordinates."""
        return self._xy

    def set_pixel_grid(self, xlim, ylim):
        """Set a pixel grid bounding box (reference tess/voronoi.py:52-73)."""
        assert len(xlim) == 2, "xlim must be (min, max) sequence"
        assert len(ylim) == 2, "ylim must be (min, max) sequence"
        self.xlim = xlim
        self.ylim = ylim
        # cached products depend on the grid
        self._segmap = None
        self._cell_areas = None

    def _grid(self):
        assert self.xlim is not None, "Need to run `set_pixel_grid()` first"
        assert self.ylim is not None, "Need to run `set_pixel_grid()` first"
        # np.arange(lo, hi) of voronoi.py:107-109: lo, lo+1, ... < hi
        W = max(int(math.ceil(self.xlim[1] - self.xlim[0])), 0)
        H = max(int(math.ceil(self.ylim[1] - self.ylim[0])), 0)
        return float(self.xlim[0]), float(self.ylim[0]), H, W

    def _segmap_device(self):
        x0, y0, H, W = self._grid()
        if H == 0 or W == 0:
            return self._eng()._empty((H, W), "int32")
        return self._eng().assign_grid(x0, y0, H, W)

    @property
    def segmap(self):
        """Segmentation map of Voronoi bin numbers for each pixel, ``segmap[y, x]``
        (reference tess/voronoi.py:75-81)."""
        if self._segmap is None:
            self._segmap = self._out(self._segmap_device(), "int64")
        return self._segmap

    def render_voronoi_field(self, nodeValues):
        """Renders the Voronoi field onto the pixel grid: ``field[y, x] = nodeValues[segmap[y, x]]``
        (reference tess/voronoi.py:83-114)."""
        assert self.xlim is not None, "Need to run `set_pixel_grid()` first"
        assert self.ylim is not None, "Need to run `set_pixel_grid()` first"
        assert len(nodeValues) == self._xy.shape[0], "Not the same number of" \
            " node values as nodes!"
        seg = self.segmap
        if self._torch_out:
            vals = nodeValues if _is_torch(nodeValues) else self._eng()._torch.as_tensor(nodeValues)
            return vals.to(seg.device)[seg]
        return np.asarray(nodeValues)[seg]

    def compute_cell_areas(self, flagmap=None):
        """Pixel count of every Voronoi cell; pixels with ``flagmap > 0`` are left out
        (reference tess/voronoi.py:116-152; the flagmap branch there writes NaN into an integer
        array and cannot run -- here it does what its docstring says)."""
        assert self._segmap is not None, "Compute a segmentation map first"
        eng = self._eng()
        seg = eng._to_dev(self._segmap, "int32")
        if flagmap is not None:
            flag = eng._to_dev(flagmap, "float64")
            seg = eng._masked_fill(seg, flag > 0, -1)
        counts = self._out(eng.label_sums(seg, None, self._xy.shape[0]), "int64")
        # np.bincount(segmap.ravel()) has length max(label) + 1
        self._cell_areas = counts[:eng._amax(seg) + 1]
        return self._cell_areas

    def partition_points(self, xy):
        """Index of the Voronoi node nearest to every point (reference tess/voronoi.py:154-173)."""
        if int(xy.shape[0]) == 0:
            return self._out(self._eng()._empty((0,), "int32"), "int64")
        return self._out(self._eng().assign_points(xy), "int64")

    def sum_cell_point_mass(self, xy, mass=None):
        """Sum of point masses within each Voronoi cell (reference tess/voronoi.py:175-196)."""
        eng = self._eng()
        idx = eng.assign_points(xy)
        sums = eng.label_sums(idx, mass, self._xy.shape[0])
        # np.bincount(idx, weights=mass): length max(idx) + 1
        return self._out(sums[:eng._amax(idx) + 1])

    def cell_point_density(self, xy, mass=None, flagmap=None):
        """Density of points in each Voronoi cell (reference tess/voronoi.py:198-223)."""
        if self._cell_areas is None:
            self.compute_cell_areas(flagmap=flagmap)
        return self.sum_cell_point_mass(xy, mass=mass) / self._cell_areas


class CVTessellation(VoronoiTessellation):
    """A centroidal Voronoi tessellation (CVT): Lloyd's algorithm on weighted points
    (reference tess/voronoi.py:226-321).

    Parameters
    ----------
    xy_points : ndarray, ``(n_points, 2)``
    dens_points : ndarray
        Density *or weight* of each point; ``(S/N)**2`` for an equal-S/N generator.
    node_xy : ndarray ``(n_nodes, 2)``
        Pre-computed generators; ``None`` makes every point its own generator (voronoi.py:298-299).
    max_iters : int
        Iteration budget of the Lloyd loop (the reference runs at most max_iters + 2 bodies).
    """
    def __init__(self, xy_points, dens_points, node_xy=None, max_iters=300, device=None, engine=None):
        self._device = device
        self._engine = engine
        self._torch_out = _is_torch(xy_points)
        xy, vbin_num = self._tessellate(xy_points, dens_points, node_xy=node_xy, max_iters=max_iters)
        super(CVTessellation, self).__init__(xy, device=device, engine=self._engine)
        self._engine.gen_owner = self       # the context already holds the converged generators
        self._vbin_num = vbin_num

    def _new_engine(self, like):
        if self._engine is None:
            dev = self._device
            if dev is None and _is_torch(like) and like.is_cuda:
                dev = like.device
            self._engine = Engine(dev)
        return self._engine

    @classmethod
    def from_image(cls, density, generators, max_iters=300, device=None, engine=None):
        """CVT of a pixel map: point coordinates are the 0-based pixel indices, non-finite pixels are
        left out, the pixel grid is set to the image (reference tess/voronoi.py:259-291).
        ``generators`` are ``(x, y)``."""
        self = cls.__new__(cls)
        self._device = device
        self._engine = engine
        self._torch_out = _is_torch(density)
        if not self._torch_out:
            density = np.asarray(density)
        _check_f64("density", density, 2)
        eng = self._new_engine(density)
        eng.set_image(density=density)
        dens_dev = eng._keep["input"][0].reshape(-1)
        good = eng._finite_mask(dens_dev)           # voronoi.py:285
        nodes, labels = self._run(eng, generators, max_iters)
        VoronoiTessellation.__init__(self, nodes, device=device, engine=eng)
        eng.gen_owner = self                # the context already holds the converged generators
        self._vbin_num = self._out(labels.reshape(-1)[good], "int64")
        self.densPoints = self._out(dens_dev[good])
        self.set_pixel_grid((0, density.shape[1]), (0, density.shape[0]))
        return self

    def _run(self, eng, node_xy, max_iters):
        if not _is_torch(node_xy):
            node_xy = np.asarray(node_xy)
        _check_f64("node_xy", node_xy, 2)
        eng.set_generators(node_xy)
        converged, _ = run_lloyd(eng, int(max_iters))
        if not converged:
            log.warning("CVT did not converge")
        self._node_weights_dev = eng.moments()[:, 1]
        return self._out(eng.generators()), eng.labels()

    def _tessellate(self, xy, densPoints, node_xy=None, max_iters=300):
        """Computes the centroidal voronoi tessellation itself (reference voronoi.py:293-305)."""
        self.densPoints = densPoints
        if not self._torch_out:
            xy, densPoints = np.asarray(xy), np.asarray(densPoints)
        _check_f64("xy", xy, 2)
        _check_f64("dens_points", densPoints, 1)
        # default generators: every point its own node (reference voronoi.py:298-299)
        if node_xy is None:
            node_xy = xy.clone() if self._torch_out else xy.copy()
        eng = self._new_engine(xy)
        eng.set_points(xy, densPoints)
        nodes, labels = self._run(eng, node_xy, max_iters)
        return nodes, self._out(labels, "int64")

    @property
    def membership(self):
        """Bin index of every (finite-density) point, int64."""
        return self._vbin_num

    @property
    def node_weights(self):
        """Weight of each Voronoi bin (sum of enclosed point masses, reference voronoi.py:312-321):
        the sum-of-weights moment of the last Lloyd iteration."""
        return self._out(self._node_weights_dev)
