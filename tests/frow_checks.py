"""Checks of the SURVEY.md section-8f helpers (tess_b200.accretion / wvt), written once and run on the
CPU SIMT-emulator engine (tests/test_frows_cpu.py) and on the GPU (tests/test_gpu_parity.py).
The expected values are literal NumPy/SciPy restatements of the reference methods cited in each."""
import numpy as np
from scipy.spatial import KDTree

from oracle import oracle as O
from conftest import gaussian_image, sampled_seeds


def _host(a):
    return a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)


def ref_centroids(seg, weights=None):
    """tess/pixel_accretion.py:212-240, statement by statement."""
    n_bins = seg.max()
    w_all = np.ones(seg.shape) if weights is None else weights
    cen = np.nan * np.empty((n_bins, 2))
    for i, b in enumerate(range(n_bins)):
        pixels = np.where(seg == b)
        if len(pixels[0]) == 0:
            continue
        w = w_all[pixels].flatten()
        fin = np.where(np.isfinite(w))[0]
        for j in (0, 1):
            cen[i, j] = np.average(pixels[j][fin], weights=w[fin])
    cen = cen[:, ::-1].copy()
    return cen[np.isfinite(cen[:, 0])]


def check_accretion_helpers(make_engine):
    from tess_b200 import accretion as A
    rng = np.random.default_rng(21)
    H, W = 37, 52
    gens = rng.uniform(0, 50, (12, 2))
    seg = O.c_assign_grid(H, W, gens).astype(np.int64)
    seg[seg == 7] = 3                                        # an empty bin in the middle
    wmap = rng.uniform(0.2, 2.0, (H, W))
    wmap[5, 6] = np.nan
    wmap[20, 30] = np.inf
    np.testing.assert_allclose(A.segmap_centroids(seg, engine=make_engine()), ref_centroids(seg), rtol=0, atol=1e-10)
    np.testing.assert_allclose(A.segmap_centroids(seg, wmap, engine=make_engine()), ref_centroids(seg, wmap), rtol=0, atol=1e-10)
    # bin S/N (pixel_accretion.py:501-517)
    img = rng.uniform(1.0, 5.0, (H, W))
    noise = rng.uniform(0.5, 1.5, (H, W))
    want = []
    for b in range(seg.max()):
        px = np.where(seg == b)
        if len(px[0]) == 0:
            continue
        want.append(np.sum(img[px].flatten()) / np.sqrt(np.sum(noise[px] ** 2.)))
    np.testing.assert_allclose(A.segmap_bin_sn(seg, img, noise, engine=make_engine()), np.array(want), rtol=1e-12)
    # cleanup (pixel_accretion.py:485-499): bin centroids in (row, col); bins 2, 3, 9 failed
    nb = seg.max() + 1
    cen_rc = np.array([[np.mean(np.where(seg == b)[0]) if (seg == b).any() else -50.0,
                        np.mean(np.where(seg == b)[1]) if (seg == b).any() else -50.0] for b in range(nb)])
    valid = np.ones(nb, dtype=bool)
    valid[[2, 3, 9]] = False
    valid[7] = False
    good = np.where(valid)[0]
    tree = KDTree(cen_rc[good])
    want_seg = seg.copy()
    for f in np.where(~valid)[0]:
        px = np.where(seg == f)
        if len(px[0]) == 0:
            continue
        _, ri = tree.query(np.vstack(px).T)
        want_seg[px] = good[ri]
    got = A.cleanup_reassign(seg, cen_rc, valid, engine=make_engine())
    assert np.array_equal(got, want_seg) and not np.isin(got, [2, 3, 9]).any()
    # point-bin nodes (point_accretion.pyx:70-84), 1-based bin numbers
    n, k = 800, 9
    xy = rng.uniform(0, 100, (n, 2))
    w = rng.uniform(0.1, 1.0, n)
    bins = rng.integers(1, k + 1, n)
    node = np.zeros((k, 2)); m = np.zeros(k)
    for i in range(n):
        j = bins[i] - 1
        node[j] += xy[i] * w[i]; m[j] += w[i]
    np.testing.assert_allclose(A.point_nodes(xy, w, bins, k, engine=make_engine()), node / m[:, None], rtol=1e-12)


def check_wvt_binning(make_engine, N=72, k=14):
    from tess_b200.wvt import wvt_binning, bin_sn
    sig = gaussian_image(N) + 1.0
    noise = np.sqrt(sig + 1.0)
    seeds = sampled_seeds((sig / noise) ** 2, k, seed=4)
    x, y = np.meshgrid(np.arange(N, dtype=float), np.arange(N, dtype=float))
    xy = np.column_stack((x.ravel(), y.ravel()))
    # (i) no scale update, (S/N)^2 weights: exactly Lloyd (tess/lloyd.pyx) with w = (S/N)^2
    r = wvt_binning(sig, noise, seeds, max_iters=4, check_every=2, tol=0.0, scale_update=False,
                    uniform_weight=False, engine=make_engine())
    rn, ri, rc, nit, _, _ = O.c_lloyd(xy, ((sig / noise) ** 2).ravel(), seeds, 4)
    assert r["n_iters"] == nit and r["converged"] == rc
    assert np.array_equal(r["labels"].ravel(), ri)
    np.testing.assert_allclose(r["nodes"], rn, rtol=0, atol=1e-9)
    lab = r["labels"].ravel()
    want_sn = np.bincount(lab, weights=sig.ravel(), minlength=k) / np.sqrt(np.bincount(lab, weights=(noise ** 2).ravel(), minlength=k))
    np.testing.assert_allclose(r["bin_sn"], want_sn, rtol=1e-12)
    # (ii) WVT: the scale update equalises the bin S/N -- the scatter must drop well below plain CVT's
    cvt = wvt_binning(sig, noise, seeds, max_iters=30, check_every=4, tol=0.0, scale_update=False,
                      uniform_weight=True, engine=make_engine())
    wvt = wvt_binning(sig, noise, seeds, max_iters=30, check_every=4, tol=0.0, scale_update=True,
                      uniform_weight=True, engine=make_engine())
    assert wvt["sn_scatter"] < 0.6 * cvt["sn_scatter"], (wvt["sn_scatter"], cvt["sn_scatter"])
    # scales are those of the last update: sqrt(area / (S/N)) of the moments one iteration earlier;
    # at least they are finite, positive and the labels match the oracle's scaled assignment
    assert np.all(np.isfinite(wvt["scales"])) and np.all(wvt["scales"] > 0)
    assert wvt["history"][0][0] == 4 and wvt["n_iters"] <= 32
    # the target is acted on: far off with this k -> not met, and the suggested k moves the median onto it
    med = wvt["median_sn"]
    off = wvt_binning(sig, noise, seeds, target_sn=3.0 * med, max_iters=12, check_every=4, engine=make_engine())
    assert off["target_met"] is False and off["k_suggested"] < k // 4
    on = wvt_binning(sig, noise, seeds, target_sn=med, max_iters=30, check_every=4, sn_tol=0.1, engine=make_engine())
    assert on["target_met"] is True and on["n_iters"] <= wvt["n_iters"] and abs(on["k_suggested"] - k) <= max(2, k // 4)
    return wvt


def check_flagmap(vt, g):
    """compute_cell_areas(flagmap) / cell_point_density(flagmap): pixels with flagmap > 0 are not
    counted (docstring of tess/voronoi.py:116-152; the reference's own branch cannot run)."""
    seg = g["segmap"]
    rng = np.random.default_rng(31)
    flag = (rng.uniform(size=seg.shape) < 0.3).astype(np.int64) * rng.integers(1, 4, seg.shape)
    flag[:3] = -1                                            # negative and zero flags keep the pixel
    keep = ~(flag > 0)
    want = np.bincount(seg[keep].ravel())
    got = vt.compute_cell_areas(flagmap=flag)
    assert np.array_equal(_host(got), want)
    vt._cell_areas = None
    dens = vt.cell_point_density(g["points"], mass=g["mass"], flagmap=flag.astype(np.float64))
    n = min(len(want), len(g["cell_mass"]))
    with np.errstate(divide="ignore", invalid="ignore"):
        np.testing.assert_allclose(_host(dens)[:n], g["cell_mass"][:n] / want[:n], rtol=1e-13)
    vt._cell_areas = None
    assert np.array_equal(_host(vt.compute_cell_areas()), np.bincount(seg.ravel()))


def check_accretor_fixture(make_engine):
    """tess_b200.accretion against OUTPUTS OF THE REFERENCE's EqualSNAccretor (tests/golden/accretor.npz,
    written by make_golden.py from the shimmed tess/pixel_accretion.py): centroids (:212-240), bin_sn
    (:501-517) and the segmentation image after cleanup() (:485-499)."""
    from tess_b200 import accretion as A
    from conftest import golden
    g = golden("accretor.npz")
    after = A.cleanup_reassign(g["seg_before"], g["bin_centroids_rc"], g["valid_bins"], engine=make_engine())
    assert np.array_equal(after, g["seg_after"])
    cen = A.segmap_centroids(g["seg_after"], g["weightmap"], engine=make_engine())
    assert cen.shape == g["centroids"].shape
    np.testing.assert_allclose(cen, g["centroids"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(A.segmap_bin_sn(g["seg_after"], g["image"], g["noise"], engine=make_engine()),
                               g["bin_sn"], rtol=1e-12)
