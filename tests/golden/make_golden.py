"""
tests/golden/make_golden.py -- generates the committed golden fixtures FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference and oracle/_ref built by
oracle/build_ref.py):      python tests/golden/make_golden.py

Everything stored here is an output of the reference's own code:
  * tess/lloyd.pyx           compiled unmodified (oracle/_ref/lloyd*.so)
  * tess/point_accretion.pyx compiled unmodified (oracle/_ref/point_accretion*.so)
  * tess/voronoi.py          imported unmodified from /root/reference with three runtime shims
                             (sys.modules['lloyd'], builtins.xrange, np.float) -- SURVEY.md 8c
  * tess/pixel_accretion.py  source text read from /root/reference, four py2->py3 textual
                             substitutions applied IN MEMORY (nothing is copied into the repo)

Per-iteration states of lloyd() are obtained from the reference by calling it with
max_iters = t-2, which executes exactly t loop bodies (tess/lloyd.pyx:85-90).

Fixtures (all tie-free by construction: integer pixel coordinates, jittered generators):
  lloyd_img64.npz     64x64 demo-style Gaussian, 11 seeds: full per-iteration trajectory
  lloyd_demo256.npz   config C1: scripts/demo_iso_sn_voronoi.py inputs, EqualSNAccretor seeds
  lloyd_points.npz    config C2 shape (small): 20000-point Gaussian starfield, EqualMassAccretor seeds
  segmap.npz          VoronoiTessellation.segmap / partition_points / sum_cell_point_mass
  from_image_nan.npz  CVTessellation.from_image on a map with NaN pixels (membership, nodes, node_weights)
  accretor.npz        the shimmed EqualSNAccretor on a 40x40 map: segmentation image before and after
                      cleanup(), _valid_bins, _current_bin_centroids, centroids, bin_sn
                      (PointAccretor.bin_nums is a private cdef attribute of the compiled reference, so
                      PointAccretor.nodes() cannot be pinned the same way)

    python tests/golden/make_golden.py [name ...]     regenerates only the named fixtures
"""
import builtins
import io
import contextlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import oracle as O  # noqa: E402


def load_reference():
    ref_lloyd = O.ref_module()                       # compiled tess/lloyd.pyx
    sys.modules["lloyd"] = ref_lloyd
    builtins.xrange = range
    if not hasattr(np, "float"):
        np.float = float
    import scipy.spatial
    import scipy.interpolate  # noqa: F401
    # tess/voronoi.py, unmodified
    vor = types.ModuleType("ref_voronoi")
    src = open(os.path.join(REF, "tess", "voronoi.py")).read()
    exec(compile(src, os.path.join(REF, "tess", "voronoi.py"), "exec"), vor.__dict__)
    # tess/pixel_accretion.py with the four textual py3 substitutions of SURVEY.md Appendix B
    src = open(os.path.join(REF, "tess", "pixel_accretion.py")).read()
    src = (src.replace(".iteritems()", ".items()").replace("xrange", "range")
              .replace("np.float)", "float)").replace("np.bool)", "bool)")
              .replace("import scipy.spatial.kdtree as kdtree", "import scipy.spatial as kdtree"))
    pa = types.ModuleType("ref_pixel_accretion")
    exec(compile(src, "pixel_accretion(py3-shimmed)", "exec"), pa.__dict__)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import point_accretion as pta                    # compiled tess/point_accretion.pyx
    return ref_lloyd, vor, pa, pta


def quiet(fn, *a, **kw):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        r = fn(*a, **kw)
    return r, buf.getvalue()


def gaussian_image(N, amp=10.0):
    """scripts/demo_iso_sn_voronoi.py:20-26 scaled to N (sigma = 50*N/256)."""
    s = 50.0 * N / 256.0
    x, y = np.meshgrid(np.arange(N, dtype=float), np.arange(N, dtype=float))
    return amp * np.exp(-((x - N / 2.0) ** 2.0 / (2.0 * s ** 2.0) + (y - N / 2.0) ** 2.0 / (2.0 * s ** 2.0)))


def sampled_seeds(w, k, seed):
    """SURVEY.md 8d seed recipe: choose k pixels with p ~ w, jitter by U(-.5,.5)."""
    rng = np.random.default_rng(seed)
    H, W = w.shape
    p = w.ravel() / w.sum()
    sel = rng.choice(H * W, k, replace=False, p=p)
    gx = (sel % W).astype(float)
    gy = (sel // W).astype(float)
    return np.column_stack((gx, gy)) + rng.uniform(-0.5, 0.5, (k, 2))


def trajectory(ref_lloyd, xy, w, seeds, max_steps):
    """Per-iteration (nodes_out, labels) from the reference via max_iters = t-2."""
    nodes, labels, conv_at = [], [], None
    for t in range(1, max_steps + 1):
        (n, i, c), _ = quiet(ref_lloyd.lloyd, xy, w, seeds, t - 2)
        nodes.append(np.asarray(n).copy())
        labels.append(np.asarray(i).astype(np.int32))
        if c:
            conv_at = t
            break
    return np.array(nodes), np.array(labels), conv_at


def make_accretor(pa):
    """EqualSNAccretor outputs (tess/pixel_accretion.py:212-240, 485-517) on a small map with a few
    non-finite pixels; the accretion loop itself is out of scope (sequential), its products are not."""
    rng = np.random.default_rng(17)
    N = 40
    img = gaussian_image(N, amp=6.0) + 0.3 + 0.05 * rng.normal(size=(N, N))
    noise = np.ones((N, N)) + 0.1 * rng.uniform(size=(N, N))
    (acc, _) = quiet(pa.EqualSNAccretor, img, noise, 12.0, start=(0, 0))
    seg_before = np.array(acc.segmap, copy=True)
    valid = np.array(acc._valid_bins, dtype=bool)
    cen_rc = np.array(acc._current_bin_centroids, dtype=float)
    quiet(acc.cleanup)
    seg_after = np.array(acc.segmap, copy=True)
    cent = np.asarray(acc.centroids, dtype=float)
    sn = np.asarray(acc.bin_sn, dtype=float)
    np.savez_compressed(os.path.join(HERE, "accretor.npz"), image=img, noise=noise, target_sn=12.0,
                        seg_before=seg_before.astype(np.int32), valid_bins=valid, bin_centroids_rc=cen_rc,
                        seg_after=seg_after.astype(np.int32), centroids=cent, bin_sn=sn,
                        weightmap=np.asarray(acc.centroid_weightmap, dtype=float))
    print("accretor: %d bins (%d failed), %d centroids" % (valid.size, int((~valid).sum()), cent.shape[0]))


def main():
    ref_lloyd, vor, pa, pta = load_reference()
    only = set(sys.argv[1:])
    if only:
        if "accretor" in only:
            make_accretor(pa)
        return

    # ------------------------------------------------------------------ lloyd_img64
    img = gaussian_image(64)
    w = (img / 1.0) ** 2.0
    seeds = sampled_seeds(w, 11, seed=7)
    xy, wv, _ = O.image_points(w)
    nodes, labels, conv_at = trajectory(ref_lloyd, xy, wv, seeds, 80)
    (fn, fi, fc), out = quiet(ref_lloyd.lloyd, xy, wv, seeds, 300)
    deltas_txt = [ln.split()[-1] for ln in out.strip().splitlines()]
    assert fc and conv_at == len(deltas_txt)
    assert np.array_equal(np.asarray(fn), nodes[-1])
    np.savez_compressed(os.path.join(HERE, "lloyd_img64.npz"), density=w, seeds=seeds,
                        nodes=nodes, labels=labels.astype(np.int16), n_iters=conv_at,
                        deltas_txt=np.array(deltas_txt))
    print("lloyd_img64: %d iterations" % conv_at)

    # ------------------------------------------------------------------ lloyd_demo256 (config C1)
    shape = (256, 256)
    img = gaussian_image(256)
    noise = np.ones(shape, dtype=float)
    pix_sn = img / noise
    (acc, _) = quiet(pa.EqualSNAccretor, img, noise, 50.0, start=(0, 0))
    quiet(acc.cleanup)
    cent = np.asarray(acc.centroids, dtype=float)
    (cvt, out) = quiet(vor.CVTessellation.from_image, pix_sn ** 2.0, cent)
    n_it = len(out.strip().splitlines())
    segmap = np.asarray(cvt.segmap)
    np.savez_compressed(os.path.join(HERE, "lloyd_demo256.npz"), seeds=cent,
                        nodes=np.asarray(cvt.nodes), membership=np.asarray(cvt.membership).astype(np.int16),
                        segmap=segmap.astype(np.int16), n_iters=n_it,
                        deltas_txt=np.array([ln.split()[-1] for ln in out.strip().splitlines()]))
    print("lloyd_demo256: %d seeds, %d iterations" % (cent.shape[0], n_it))

    # ------------------------------------------------------------------ lloyd_points (config C2 shape, small)
    rng = np.random.default_rng(1)
    n = 20000
    x = rng.normal(250.0, 100.0, n)
    y = rng.normal(250.0, 50.0, n)
    ok = (x > 0) & (y > 0) & (x < 500) & (y < 500)
    pxy = np.column_stack((x[ok], y[ok]))
    pw = rng.uniform(0.1, 1.0, pxy.shape[0])
    accp = pta.EqualMassAccretor(pxy, pw, 20.0)
    quiet(accp.accrete)
    (pseeds, _) = quiet(accp.nodes)
    pseeds = np.asarray(pseeds, dtype=float)
    (cvtp, out) = quiet(vor.CVTessellation, pxy, pw, node_xy=pseeds, max_iters=48)
    n_itp = len(out.strip().splitlines())
    np.savez_compressed(os.path.join(HERE, "lloyd_points.npz"), xy=pxy, w=pw, seeds=pseeds,
                        nodes=np.asarray(cvtp.nodes), membership=np.asarray(cvtp.membership).astype(np.int32),
                        node_weights=np.asarray(cvtp.node_weights), n_iters=n_itp, max_iters=48)
    print("lloyd_points: %d points, %d seeds, %d iterations" % (pxy.shape[0], pseeds.shape[0], n_itp))

    # ------------------------------------------------------------------ segmap / partition (voronoi.py:75-196)
    rng = np.random.default_rng(11)
    gens = np.column_stack((rng.uniform(-5, 110, 40), rng.uniform(0, 90, 40)))
    vt = vor.VoronoiTessellation(gens)
    vt.set_pixel_grid((3, 99), (-2, 78))
    seg = np.asarray(vt.segmap)                       # shape (ylen, xlen), float64 under scipy>=1.x
    vals = rng.normal(size=40)
    field = np.asarray(vt.render_voronoi_field(vals))
    pts = np.column_stack((rng.uniform(0, 100, 500), rng.uniform(0, 80, 500)))
    mass = rng.uniform(0.5, 2.0, 500)
    part = np.asarray(vt.partition_points(pts))
    cmass = np.asarray(vt.sum_cell_point_mass(pts, mass=mass))
    np.savez_compressed(os.path.join(HERE, "segmap.npz"), gens=gens, xlim=np.array([3, 99]),
                        ylim=np.array([-2, 78]), segmap=seg.astype(np.int32), values=vals, field=field,
                        points=pts, mass=mass, partition=part.astype(np.int32), cell_mass=cmass)
    print("segmap:", seg.shape, seg.dtype)

    # ------------------------------------------------------------------ from_image with NaN pixels
    rng = np.random.default_rng(2)
    img = gaussian_image(96) + 1.0
    nz = np.sqrt(img + 1.0)
    dens = (img / nz) ** 2.0
    bad = rng.choice(96 * 96, 96 * 96 // 100, replace=False)
    dens.ravel()[bad] = np.nan
    gseeds = sampled_seeds(np.nan_to_num(dens), 30, seed=5)
    (cv, out) = quiet(vor.CVTessellation.from_image, dens, gseeds, max_iters=300)
    np.savez_compressed(os.path.join(HERE, "from_image_nan.npz"), density=dens, seeds=gseeds,
                        nodes=np.asarray(cv.nodes), membership=np.asarray(cv.membership).astype(np.int16),
                        node_weights=np.asarray(cv.node_weights), segmap=np.asarray(cv.segmap).astype(np.int16),
                        n_iters=len(out.strip().splitlines()))
    print("from_image_nan: %d iterations" % len(out.strip().splitlines()))

    make_accretor(pa)


if __name__ == "__main__":
    main()
