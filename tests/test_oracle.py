"""
The oracle (oracle/) is pinned against (i) the committed golden fixtures, which are outputs of the
reference's own code (tests/golden/make_golden.py), and (ii) the compiled reference itself
(oracle/_ref, tess/lloyd.pyx built unmodified) when it is present.  CPU only.
"""
import numpy as np
import pytest

from conftest import golden, gaussian_image, sampled_seeds
from oracle import oracle as O


def test_c_and_numpy_restatements_agree_bitwise():
    rng = np.random.default_rng(3)
    xy = rng.uniform(0, 50, (3000, 2))
    node = rng.uniform(0, 50, (40, 2))
    w = rng.uniform(0.1, 2, 3000)
    a = O.c_assign_points(xy, node)
    b = O.np_assign_points(xy, node)
    assert np.array_equal(a, b)
    assert np.array_equal(O.c_accumulate(xy, w, a, 40), O.np_moments(xy, w, b, 40))
    sc = rng.uniform(0.5, 2.0, 40)
    assert np.array_equal(O.c_assign_points(xy, node, sc), O.np_assign_points(xy, node, sc))
    # scale == 1 degenerates bit-exactly to the unscaled rule
    assert np.array_equal(O.c_assign_points(xy, node, np.ones(40)), a)


def test_restatement_matches_ckdtree_on_tie_free_input():
    rng = np.random.default_rng(4)
    node = rng.uniform(0, 64, (100, 2))
    lab = O.c_assign_grid(64, 64, node)
    x, y = np.meshgrid(np.arange(64.0), np.arange(64.0))
    xy = np.column_stack((x.ravel(), y.ravel()))
    assert np.array_equal(lab.ravel(), O.kd_assign(xy, node))


def test_golden_img64_trajectory():
    """Every iteration of the reference's lloyd() on the 64x64 fixture: labels and nodes bitwise."""
    g = golden("lloyd_img64.npz")
    xy, w, _ = O.image_points(g["density"])
    rec = []
    node, idx, conv = O.np_lloyd(xy, w, g["seeds"], 300, record=rec)
    assert conv and len(rec) == int(g["n_iters"])
    for t, r in enumerate(rec):
        assert np.array_equal(r["labels"], g["labels"][t].astype(np.int64)), t
        assert np.array_equal(r["nodes_out"], g["nodes"][t]), t
        assert "%.2e" % r["delta"] == str(g["deltas_txt"][t]), t
    # the C restatement follows the same trajectory
    cn, ci, cc, nit, deltas, _ = O.c_lloyd(xy, w, g["seeds"], 300)
    assert cc and nit == int(g["n_iters"])
    assert np.array_equal(cn, g["nodes"][-1]) and np.array_equal(ci, g["labels"][-1].astype(np.int64))


def test_golden_demo256_final_state():
    """Config C1 (scripts/demo_iso_sn_voronoi.py): converged nodes / membership / segmap."""
    g = golden("lloyd_demo256.npz")
    w = gaussian_image(256) ** 2.0
    xy, wv, _ = O.image_points(w)
    cn, ci, cc, nit, deltas, _ = O.c_lloyd(xy, wv, g["seeds"], 300)
    assert cc and nit == int(g["n_iters"])
    assert np.array_equal(cn, g["nodes"])
    assert np.array_equal(ci, g["membership"].astype(np.int64))
    assert np.array_equal(O.c_assign_grid(256, 256, cn), g["segmap"].astype(np.int64))
    assert ["%.2e" % d for d in deltas] == [str(s) for s in g["deltas_txt"]]


def test_golden_points():
    g = golden("lloyd_points.npz")
    cn, ci, cc, nit, _, mom = O.c_lloyd(g["xy"], g["w"], g["seeds"], int(g["max_iters"]))
    assert nit == int(g["n_iters"]) and not cc
    assert np.array_equal(cn, g["nodes"]) and np.array_equal(ci, g["membership"].astype(np.int64))
    np.testing.assert_allclose(mom[:, 1], g["node_weights"], rtol=1e-13)


def test_golden_segmap_and_partition():
    g = golden("segmap.npz")
    H = int(g["ylim"][1] - g["ylim"][0]); W = int(g["xlim"][1] - g["xlim"][0])
    lab = O.c_assign_grid(H, W, g["gens"], x0=float(g["xlim"][0]), y0=float(g["ylim"][0]))
    assert np.array_equal(lab, g["segmap"].astype(np.int64))
    assert np.array_equal(O.c_assign_points(g["points"], g["gens"]), g["partition"].astype(np.int64))
    assert np.array_equal(g["values"][lab], g["field"])


def test_golden_from_image_nan():
    g = golden("from_image_nan.npz")
    xy, w, good = O.image_points(g["density"])
    assert good.size == g["membership"].size
    cn, ci, cc, nit, _, mom = O.c_lloyd(xy, w, g["seeds"], 300)
    assert nit == int(g["n_iters"])
    assert np.array_equal(cn, g["nodes"]) and np.array_equal(ci, g["membership"].astype(np.int64))
    np.testing.assert_allclose(mom[:, 1], g["node_weights"], rtol=1e-13)


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref (compiled reference) not built")
def test_restatement_vs_compiled_reference():
    """The compiled, unmodified tess/lloyd.pyx against the restatement: bitwise, incl. stop rule."""
    w = gaussian_image(48) ** 2.0
    seeds = sampled_seeds(w, 9, seed=21)
    xy, wv, _ = O.image_points(w)
    for max_iters in (0, 3, 300):
        rn, ri, rc, out = O.ref_lloyd(xy, wv, seeds, max_iters, want_stdout=True)
        cn, ci, cc, nit, deltas, _ = O.c_lloyd(xy, wv, seeds, max_iters)
        assert rc == cc and nit == len(out.strip().splitlines())
        assert np.array_equal(rn, cn) and np.array_equal(ri, ci)
        assert out.strip().splitlines() == ["CVT Delta %03d %.2e" % (i, d) for i, d in enumerate(deltas)]


def test_empty_bin_goes_to_origin():
    """lloyd.pyx:68-74: a generator that owns no point is parked at (0, 0)."""
    xy = np.array([[5.0, 5.0], [6.0, 5.0], [5.0, 6.0]])
    w = np.ones(3)
    node = np.array([[5.2, 5.1], [100.0, 100.0]])
    idx = O.c_assign_points(xy, node)
    mom = O.c_accumulate(xy, w, idx, 2)
    new, delta = O.c_update_nodes(mom, node)
    assert np.array_equal(new[1], [0.0, 0.0]) and mom[1, 0] == 0
