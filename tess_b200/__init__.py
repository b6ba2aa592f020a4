"""
tess_b200 -- B200 (sm_100a) implementation of the Lloyd / centroidal-Voronoi hot path of
jonathansick/tess behind tess's own Python API:

    tess_b200.lloyd.lloyd                 <->  tess.lloyd.lloyd              (tess/lloyd.pyx:10-90)
    tess_b200.voronoi.VoronoiTessellation <->  tess.voronoi.VoronoiTessellation (tess/voronoi.py:31-223)
    tess_b200.voronoi.CVTessellation      <->  tess.voronoi.CVTessellation      (tess/voronoi.py:226-321)
    tess_b200.accretion                   <->  accretor centroids / bin_sn / cleanup (tess/pixel_accretion.py:212-240,485-517)
    tess_b200.wvt.wvt_binning                  Lloyd/WVT to a target S/N on the device
    tess_b200.fitsio                           single-HDU FITS maps in / segmaps out (astropy is absent)

Everything numeric runs in hand-written CUDA kernels inside libtess_b200.so (built by
`python -m tess_b200.build`, C ABI in include/tess_b200.h); Python is a thin ctypes host layer.
There is no CPU fallback: the compute entry points raise if the library or a GPU is missing.
"""
__version__ = "0.1.0"

from . import _capi  # noqa: F401
from ._capi import TessError, TessNonFinite  # noqa: F401


def __getattr__(name):
    # lazy: importing the package must work on a GPU-less box (build checks, CPU tests)
    if name in ("lloyd", "voronoi", "engine", "sharded", "accretion", "wvt", "fitsio"):
        import importlib
        return importlib.import_module("." + name, __name__)
    if name in ("VoronoiTessellation", "CVTessellation"):
        from . import voronoi
        return getattr(voronoi, name)
    if name == "Engine":
        from .engine import Engine
        return Engine
    raise AttributeError(name)
