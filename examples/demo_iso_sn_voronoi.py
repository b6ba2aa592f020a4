"""examples/demo_iso_sn_voronoi.py -- the reference's scripts/demo_iso_sn_voronoi.py (config C1) with
`tess` replaced by `tess_b200`: the 256^2 Gaussian image with constant noise (:20-26), generators
from the equal-S/N accretor, `CVTessellation.from_image(pix_sn ** 2, generators)` (:38) -- the same
call, now on the GPU -- and the two segmentation maps written as FITS instead of matplotlib figures
(neither matplotlib nor astropy is in this image).

The accretor (`EqualSNAccretor(img, noise, 50., start=(0, 0))`, :29-31) is a sequential CPU seed
generator and out of scope; its 271 centroids for exactly these inputs are in
tests/golden/lloyd_demo256.npz, made by running the reference itself (tests/golden/make_golden.py).

    python examples/demo_iso_sn_voronoi.py [outdir]
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
    shape = (256, 256)
    pix_x, pix_y = np.meshgrid(np.arange(shape[1], dtype=float), np.arange(shape[0], dtype=float))
    img = 10. * np.exp(-((pix_x - 128.) ** 2. / (2. * 50. ** 2.) + (pix_y - 128.) ** 2. / (2. * 50. ** 2.)))
    noise = np.ones(shape, dtype=float)
    pix_sn = img / noise
    generator_centroids = np.load(os.path.join(ROOT, "tests", "golden", "lloyd_demo256.npz"))["seeds"]

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
