"""examples/wvt_demo.py -- config C1 end to end on the GPU, from files to files: the 256^2 Gaussian
S/N map of the reference's scripts/demo_iso_sn_voronoi.py:20-26, seeds drawn with p ~ (S/N)^2
(the reference takes them from its CPU accretor), WVT iteration to equal S/N, segmentation map
written as FITS (as tess/tests/test_pixel_accretion.py:35 does with astropy).

    python examples/wvt_demo.py [N] [k] [outdir]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tess_b200.fitsio import read_image, write_image   # noqa: E402
from tess_b200.wvt import wvt_binning                  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
k = int(sys.argv[2]) if len(sys.argv) > 2 else 271
out = sys.argv[3] if len(sys.argv) > 3 else "."
s = 50.0 * N / 256.0
x, y = np.meshgrid(np.arange(N, dtype=float), np.arange(N, dtype=float))
signal = 10.0 * np.exp(-((x - N / 2.0) ** 2 + (y - N / 2.0) ** 2) / (2.0 * s ** 2)) + 1.0
noise = np.sqrt(signal + 1.0)
write_image(os.path.join(out, "signal.fits"), signal)
write_image(os.path.join(out, "noise.fits"), noise)
signal, _ = read_image(os.path.join(out, "signal.fits"))
noise, _ = read_image(os.path.join(out, "noise.fits"))
rng = np.random.default_rng(0)
w = (signal / noise) ** 2
sel = rng.choice(N * N, k, replace=False, p=w.ravel() / w.sum())
seeds = np.column_stack((sel % N, sel // N)).astype(float) + rng.uniform(-0.5, 0.5, (k, 2))
r = wvt_binning(signal, noise, seeds, max_iters=100, check_every=5, tol=1e-3, verbose=True)
write_image(os.path.join(out, "segmap.fits"), r["labels"].astype(np.int32),
            header={"NBINS": k, "NITER": r["n_iters"], "SNSCAT": r["sn_scatter"]})
print("%d bins, %d iterations, S/N = %.2f +- %.1f %%, segmap -> %s"
      % (k, r["n_iters"], np.nanmean(r["bin_sn"]), 100 * r["sn_scatter"], os.path.join(out, "segmap.fits")))
