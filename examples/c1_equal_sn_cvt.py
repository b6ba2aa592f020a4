"""examples/c1_equal_sn_cvt.py -- what the reference's scripts/demo_iso_sn_voronoi.py does (config C1), with
`tess` replaced by `tess_b200`: the 256^2 Gaussian image with constant noise (:20-26), generators
from the equal-S/N accretor, `CVTessellation.from_image(pix_sn ** 2, generators)` (:38) -- the same
call, now on the GPU -- and the two segmentation maps written as FITS instead of matplotlib figures
(neither matplotlib nor astropy is in this image).

The accretor (`EqualSNAccretor(img, noise, 50., start=(0, 0))`, :29-31) is a sequential CPU seed
generator and out of scope; its 271 centroids for exactly these inputs are in
tests/golden/lloyd_demo256.npz, made by running the reference itself (tests/golden/make_golden.py).

    python examples/c1_equal_sn_cvt.py [outdir]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tess_b200.accretion import segmap_bin_sn      # noqa: E402
from tess_b200.fitsio import write_image           # noqa: E402
from tess_b200.voronoi import CVTessellation       # noqa: E402


def main(out="."):
    n, sigma, amplitude = 256, 50.0, 10.0
    col, row = np.meshgrid(np.arange(n, dtype=np.float64), np.arange(n, dtype=np.float64))
    two_var = 2.0 * sigma ** 2.0
    gx, gy = (col - n / 2.0) ** 2.0 / two_var, (row - n / 2.0) ** 2.0 / two_var
    img = amplitude * np.exp(-(gx + gy))                        # the script's 2-D Gaussian (:24-25), bitwise
    noise = np.ones_like(img)                                   # constant noise, :26
    pix_sn = img / noise
    fixture = np.load(os.path.join(ROOT, "tests", "golden", "lloyd_demo256.npz"))
    generator_centroids = fixture["seeds"]                      # EqualSNAccretor(img, noise, 50.) centroids

    cvt = CVTessellation.from_image(pix_sn ** 2., generator_centroids)
    cvt_xy = cvt.nodes
    segimage = cvt.segmap
    sn = segmap_bin_sn(segimage, img, noise)
    write_image(os.path.join(out, "iso_sn_voronoi.fits"), segimage.astype(np.int32),
                header={"NBINS": int(cvt_xy.shape[0])})
    print("%d Voronoi bins; S/N per bin: median %.1f, 16-84 %% %.1f-%.1f; segmap -> %s"
          % (cvt_xy.shape[0], np.median(sn), np.percentile(sn, 16), np.percentile(sn, 84),
             os.path.join(out, "iso_sn_voronoi.fits")))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else ".")
