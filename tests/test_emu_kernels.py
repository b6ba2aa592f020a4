"""
Kernel LOGIC on the CPU: tess_b200/csrc compiled against the SIMT emulator (tests/emu) and driven
through the product's own host code (tess_b200.engine / lloyd / voronoi) with NumPy memory.
This is not a product path (the product has no CPU build); it catches indexing, warp-choreography
and latch errors on the GPU-less build box.  The same checks run on the real GPU in
tests/test_gpu_parity.py.
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, golden, gaussian_image, sampled_seeds
from oracle import oracle as O

sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
from harness import EmuEngine  # noqa: E402


@pytest.fixture()
def eng():
    e = EmuEngine()
    yield e
    e.close()


def test_img64_trajectory_matches_reference(eng):
    g = golden("lloyd_img64.npz")
    eng.set_image(density=g["density"])
    eng.set_generators(g["seeds"])
    T = int(g["n_iters"])
    for t in range(T):
        eng.step(1)
        assert np.array_equal(eng.labels().ravel(), g["labels"][t].astype(np.int32)), t
        assert np.abs(eng.generators() - g["nodes"][t]).max() < 1e-9, t
        delta, nch, it = eng.stats()
        assert it == t + 1
        assert abs(delta - float(g["deltas_txt"][t])) <= 6e-3 * max(delta, 1e-300) or (delta == 0.0 and t == T - 1)
    assert eng.stats()[0] == 0.0 and eng.status() == (True, False)
    eng.step(3)                                   # latch closed: further steps are no-ops
    assert eng.stats()[2] == T


def test_stop_rule(eng):
    g = golden("lloyd_img64.npz")
    eng.set_image(density=g["density"])
    for max_iters, want in ((300, (True, int(g["n_iters"]))), (3, (False, 5)), (0, (False, 2)), (-5, (False, 1))):
        eng.set_generators(g["seeds"])
        assert eng.run(max_iters) == want


def test_moments_and_sn_mode(eng):
    rng = np.random.default_rng(5)
    sig = gaussian_image(80) + 1.0
    noise = np.sqrt(sig + 1.0)
    sig[3, 7] = np.nan
    noise[40, 41] = np.inf
    w = (sig / noise) ** 2
    seeds = sampled_seeds(np.nan_to_num(w, posinf=0.0), 14, seed=8)
    eng.set_image(signal=sig, noise=noise)
    eng.set_generators(seeds)
    eng.accumulate()
    lab = eng.labels()
    assert np.array_equal(lab, O.c_assign_grid(80, 80, seeds))        # masked pixels still get a label
    good = np.isfinite(w) & np.isfinite(sig) & np.isfinite(noise)
    x, y = np.meshgrid(np.arange(80.0), np.arange(80.0))
    xy = np.column_stack((x.ravel(), y.ravel()))
    idx = np.where(good.ravel(), lab.ravel(), -1).astype(np.int64)
    ref = O.np_moments(xy, np.nan_to_num(w.ravel(), posinf=0.0), idx, 14, signal=np.nan_to_num(sig.ravel()),
                       noise=np.nan_to_num(noise.ravel(), posinf=0.0))
    np.testing.assert_allclose(eng.moments(), ref, rtol=1e-12, atol=0)
    eng.update()
    new, _ = O.np_update_nodes(ref, seeds)
    np.testing.assert_allclose(eng.generators(), new, rtol=0, atol=1e-9)


def test_wvt_scaled_assignment(eng):
    rng = np.random.default_rng(6)
    node = rng.uniform(0, 100, (60, 2))
    scale = rng.uniform(0.5, 3.0, 60)
    eng.set_generators(node, scale)
    lab = eng.assign_grid(-3.0, 2.0, 70, 90)
    assert np.array_equal(lab, O.c_assign_grid(70, 90, node, x0=-3.0, y0=2.0, scale=scale))
    assert np.array_equal(lab, eng.assign_grid(-3.0, 2.0, 70, 90, exhaustive=True))
    pts = rng.uniform(-20, 120, (700, 2))
    assert np.array_equal(eng.assign_points(pts), O.c_assign_points(pts, node, scale))


def test_ties_go_to_lowest_index(eng):
    # generators on half-integers: thousands of exact ties on an integer pixel grid
    gx, gy = np.meshgrid(np.arange(0.5, 40, 4.0), np.arange(0.5, 40, 4.0))
    node = np.column_stack((gx.ravel(), gy.ravel()))
    node = node[np.random.default_rng(1).permutation(node.shape[0])]
    eng.set_generators(node)
    assert np.array_equal(eng.assign_grid(0, 0, 40, 40), O.c_assign_grid(40, 40, node))


def test_far_generators_and_odd_shapes(eng):
    # all generators in one corner, odd width (scalar path), window offset from the generators
    rng = np.random.default_rng(9)
    node = rng.uniform(0, 6, (25, 2))
    eng.set_generators(node)
    for (x0, y0, H, W) in ((0, 0, 130, 67), (50.5, -20.25, 65, 129), (-300, 400, 3, 5)):
        assert np.array_equal(eng.assign_grid(x0, y0, H, W), O.c_assign_grid(H, W, node, x0=x0, y0=y0)), (x0, y0)
    # a single generator; more generators than the per-tile list can hold (exhaustive-scan fallback)
    eng.set_generators(node[:1])
    assert (eng.assign_grid(0, 0, 70, 70) == 0).all()
    dense = rng.uniform(0, 40, (1300, 2))
    eng.set_generators(dense)
    assert np.array_equal(eng.assign_grid(0, 0, 40, 40), O.c_assign_grid(40, 40, dense))


def test_voronoi_api_matches_reference_fixture():
    from tess_b200.voronoi import VoronoiTessellation
    g = golden("segmap.npz")
    vt = VoronoiTessellation(g["gens"], engine=EmuEngine())
    vt.set_pixel_grid(tuple(g["xlim"]), tuple(g["ylim"]))
    assert np.array_equal(vt.segmap, g["segmap"]) and vt.segmap.dtype == np.int64
    assert np.array_equal(vt.render_voronoi_field(g["values"]), g["field"])
    assert np.array_equal(vt.partition_points(g["points"]), g["partition"])
    np.testing.assert_allclose(vt.sum_cell_point_mass(g["points"], mass=g["mass"]), g["cell_mass"], rtol=1e-13)
    assert np.array_equal(vt.compute_cell_areas(), np.bincount(g["segmap"].ravel()))
    flag = np.zeros(g["segmap"].shape); flag[:10] = 1
    assert np.array_equal(vt.compute_cell_areas(flag)[:5], np.bincount(g["segmap"][10:].ravel(), minlength=40)[:5])
    import frow_checks
    frow_checks.check_flagmap(vt, g)
    # a second object on the same Engine replaces the context's generators: the first one notices
    vt2 = VoronoiTessellation(g["gens"][:7] + 0.25, engine=vt._engine)
    vt2.set_pixel_grid((0, 20), (0, 20))
    assert np.array_equal(vt2.segmap, O.c_assign_grid(20, 20, g["gens"][:7] + 0.25))
    assert np.array_equal(vt.partition_points(g["points"]), g["partition"])


def test_from_image_with_nan_matches_reference_fixture():
    from tess_b200.voronoi import CVTessellation
    g = golden("from_image_nan.npz")
    cv = CVTessellation.from_image(g["density"], g["seeds"], engine=EmuEngine())
    assert np.abs(cv.nodes - g["nodes"]).max() < 1e-6
    assert np.array_equal(cv.membership, g["membership"])
    np.testing.assert_allclose(cv.node_weights, g["node_weights"], rtol=1e-12)
    assert np.array_equal(cv.segmap, g["segmap"])
    assert (cv.xlim, cv.ylim) == ((0, 96), (0, 96))


def test_lloyd_points_small(eng):
    from tess_b200.lloyd import lloyd
    rng = np.random.default_rng(12)
    xy = np.column_stack((rng.normal(250, 100, 1500), rng.normal(250, 50, 1500)))
    w = rng.uniform(0.1, 1.0, 1500)
    seeds = xy[rng.choice(1500, 40, replace=False)] + rng.uniform(-0.1, 0.1, (40, 2))
    n, i, c = lloyd(xy, w, seeds, 6, engine=eng)
    rn, ri, rc, nit, _, _ = O.c_lloyd(xy, w, seeds, 6)
    assert c == rc and np.array_equal(i, ri) and i.dtype == np.int64
    np.testing.assert_allclose(n, rn, rtol=0, atol=1e-9)
    with pytest.raises(ValueError):
        lloyd(xy.astype(np.float32), w, seeds, 3, engine=eng)       # reference: Buffer dtype mismatch


def test_zero_weight_bin_raises_like_the_reference(eng):
    from tess_b200.lloyd import lloyd
    xy = np.array([[0.0, 0.0], [1.0, 0.0], [10.0, 0.0], [11.0, 0.0]])
    w = np.array([1.0, 1.0, 1.0, -1.0])          # second bin: count 2, sum w == 0 -> NaN node
    with pytest.raises(ValueError):
        lloyd(xy, w, np.array([[0.4, 0.1], [10.4, 0.1]]), 10, engine=eng)


def test_options_and_dense_fallbacks(eng):
    """WVT scale update, uniform weights, list diagnostics and the scan fallbacks (see the GPU twin in
    tests/test_gpu_parity.py)."""
    from tess_b200 import _capi
    N, k = 96, 40
    sig = gaussian_image(N) + 1.0
    noise = np.sqrt(sig + 1.0)
    seeds = sampled_seeds((sig / noise) ** 2, k, seed=6)
    x, y = np.meshgrid(np.arange(N, dtype=float), np.arange(N, dtype=float))
    xy = np.column_stack((x.ravel(), y.ravel()))
    eng.set_option(_capi.OPT_UNIFORM_WEIGHT, 1)
    eng.set_option(_capi.OPT_WVT_SCALE_UPDATE, 1)
    eng.set_image(signal=sig, noise=noise)
    eng.set_generators(seeds)
    eng.step(1)
    lab = O.c_assign_grid(N, N, seeds)
    assert np.array_equal(eng.labels(), lab)
    rm = O.np_moments(xy, np.ones(N * N), lab.ravel(), k, signal=sig.ravel(), noise=noise.ravel())
    np.testing.assert_allclose(eng.moments(), rm, rtol=1e-12, atol=0)
    sc = np.sqrt(rm[:, 0] / O.bin_sn(rm))
    np.testing.assert_allclose(eng.scales(), sc, rtol=1e-13)
    node, _ = O.np_update_nodes(rm, seeds)
    eng.step(1)
    assert np.array_equal(eng.labels(), O.c_assign_grid(N, N, node, scale=sc))
    eng.set_option(_capi.OPT_UNIFORM_WEIGHT, 0)
    eng.set_option(_capi.OPT_WVT_SCALE_UPDATE, 0)
    # dense generators: quadrant lists beyond the fast path, tile overflow
    rng = np.random.default_rng(3)
    dense = rng.uniform(0, 64, (64 * 64 // 6, 2))
    w = rng.uniform(0.5, 1.5, (64, 64))
    eng.set_image(density=w)
    eng.set_generators(dense)
    eng.accumulate()
    ref = O.c_assign_grid(64, 64, dense)
    assert np.array_equal(eng.labels(), ref)
    x, y = np.meshgrid(np.arange(64.0), np.arange(64.0))
    rm = O.np_moments(np.column_stack((x.ravel(), y.ravel())), w.ravel(), ref.ravel(), dense.shape[0])
    np.testing.assert_allclose(eng.moments()[:, :4], rm[:, :4], rtol=1e-12, atol=0)
    st = eng.list_stats()
    assert st["ring_fallback"] + st["scan_list"] > 0


def test_every_pixel_its_own_generator_and_dense_scaled(eng):
    """Bins of one pixel (the reference's default when no node_xy is given, voronoi.py:298-299) and a
    dense scaled table: every tile overflows the list builder and runs the per-pixel ring search over
    a refined generator grid."""
    H, W = 70, 45
    x, y = np.meshgrid(np.arange(W, dtype=float), np.arange(H, dtype=float))
    xy = np.column_stack((x.ravel(), y.ravel()))
    w = np.random.default_rng(0).uniform(0.5, 1.5, (H, W))
    eng.set_image(density=w)
    eng.set_generators(xy)
    eng.accumulate()
    assert np.array_equal(eng.labels().ravel(), np.arange(H * W, dtype=np.int32))
    m = eng.moments()
    np.testing.assert_allclose(m[:, 0], 1.0)
    np.testing.assert_allclose(m[:, 1], w.ravel(), rtol=1e-15)
    assert eng.list_stats()["ring_fallback"] > 0
    eng.update()
    np.testing.assert_allclose(eng.generators(), xy, rtol=0, atol=1e-12)
    # dense + scaled metric
    rng = np.random.default_rng(1)
    node = rng.uniform(0, 45, (700, 2))
    scale = rng.uniform(0.7, 1.6, 700)
    eng.set_generators(node, scale=scale)
    eng.accumulate()
    assert np.array_equal(eng.labels(), O.c_assign_grid(H, W, node, scale=scale))


def test_lloyd_edge_cases_follow_the_reference(eng):
    """Probed on the compiled reference (oracle/_ref): n = 0 returns zero nodes and converged; a
    longer w is read partially, a shorter one is an IndexError; k = 0, wrong widths, wrong dtypes and
    non-finite point coordinates are ValueErrors (typed memoryviews / cKDTree)."""
    from tess_b200.lloyd import lloyd
    g = np.array([[1., 2.], [3., 4.]])
    n, i, c = lloyd(np.empty((0, 2)), np.empty(0), g, 10, engine=eng)
    assert np.array_equal(n, np.zeros((2, 2))) and i.shape == (0,) and i.dtype == np.int64 and c is True
    xy = np.random.default_rng(0).uniform(0, 5, (10, 2))
    with pytest.raises(IndexError):
        lloyd(xy, np.ones(5), g, 10, engine=eng)
    a = lloyd(xy, np.ones(15), g, 10, engine=eng)
    b = O.c_lloyd(xy, np.ones(10), g, 10)
    np.testing.assert_allclose(a[0], b[0], atol=1e-12)
    assert np.array_equal(a[1], b[1]) and a[2] == b[2]
    for bad in ((xy, np.ones(10), np.empty((0, 2))), (np.zeros((10, 3)), np.ones(10), g),
                (xy, np.ones(10), np.zeros((2, 3))), (np.full((3, 2), np.nan), np.ones(3), g),
                (xy, np.ones(10, dtype=int), g)):
        with pytest.raises(ValueError):
            lloyd(*bad, 10, engine=eng)
    from tess_b200.voronoi import VoronoiTessellation
    vt = VoronoiTessellation(g, engine=eng)
    with pytest.raises(ValueError):                      # scipy's KDTree.query: 'x' must be finite
        vt.partition_points(np.array([[0.5, np.inf]]))


def test_degenerate_point_sets(eng):
    """Point-query grids for degenerate geometry: points on a line, a single point, generators far
    from the points (disjoint bounding boxes), one far outlier generator.  The grid covers the overlap
    of the two bounding boxes and everything else is clamped into border cells (exact)."""
    from tess_b200.lloyd import lloyd
    from tess_b200.voronoi import VoronoiTessellation
    rng = np.random.default_rng(9)

    def brute(pts, g):
        d = (pts[:, None, 0] - g[None, :, 0]) ** 2 + (pts[:, None, 1] - g[None, :, 1]) ** 2
        return d.argmin(axis=1)

    g = rng.uniform(0, 100, (200, 2))
    cases = {
        "vertical line": np.column_stack((np.full(300, 37.25), rng.uniform(-20, 120, 300))),
        "horizontal line": np.column_stack((rng.uniform(0, 100, 300), np.full(300, 1e-3))),
        "single point": np.array([[50.5, 49.5]]),
        "identical points": np.tile([[10.0, 90.0]], (50, 1)),
        "disjoint boxes": rng.uniform(1000, 1100, (300, 2)),
        "points around": rng.uniform(-500, 600, (300, 2)),
    }
    vt = VoronoiTessellation(g, engine=eng)
    for name, pts in cases.items():
        assert np.array_equal(vt.partition_points(pts), brute(pts, g)), name
    g2 = g.copy()
    g2[0] = [1e7, -3e6]                                   # one far outlier generator
    vt2 = VoronoiTessellation(g2, engine=eng)
    pts = rng.uniform(0, 100, (400, 2))
    assert np.array_equal(vt2.partition_points(pts), brute(pts, g2))
    # and a Lloyd run on a line of points (the reference's kd-tree has no trouble with it)
    pts = cases["vertical line"]
    w = rng.uniform(0.5, 1.0, 300)
    seeds = pts[rng.choice(300, 12, replace=False)] + rng.uniform(-0.1, 0.1, (12, 2))
    n, i, c = lloyd(pts, w, seeds, 20, engine=eng)
    rn, ri, rc = O.c_lloyd(pts, w, seeds, 20)[:3]
    assert np.array_equal(i, ri) and c == rc
    np.testing.assert_allclose(n, rn, rtol=0, atol=1e-9)


def test_randomised_small_maps_against_the_oracle(eng):
    """The CPU twin of tools/fuzz_parity.py at emulator sizes: random shapes, window origins, bin sizes
    from thousands of pixels down to one, ties, duplicates, scales, all kernel modes -- labels
    bit-exact and moments to 1e-12 against the oracle."""
    rng = np.random.default_rng(17)
    for case in range(14):
        H, W = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        x0, y0 = (float(rng.uniform(-20, 20)), float(rng.uniform(-20, 20))) if case % 2 else (0.0, 0.0)
        area = float(np.exp(rng.uniform(np.log(0.8), np.log(3000.0))))
        k = int(min(max(1, H * W / area), 4000))
        g = np.column_stack((rng.uniform(x0 - 5, x0 + W + 5, k), rng.uniform(y0 - 5, y0 + H + 5, k)))
        if case % 3 == 0:
            m = max(1, k // 8)
            g[:m] = np.round(g[:m] - [x0, y0]) + [x0, y0]      # on pixel centres: exact ties
        if k > 3:
            g[k // 2] = g[1]
        scale = rng.uniform(0.6, 1.7, k) if case % 4 == 1 else None
        w = rng.uniform(0.5, 1.5, (H, W))
        if case % 5 == 2:
            w[rng.integers(0, H), rng.integers(0, W)] = np.nan
        x, y = np.meshgrid(x0 + np.arange(W, dtype=float), y0 + np.arange(H, dtype=float))
        pts = np.column_stack((x.ravel(), y.ravel()))
        ref = O.c_assign_points(pts, g, scale=scale).reshape(H, W)
        eng.set_image(density=w, x0=x0, y0=y0)
        eng.set_generators(g, scale=scale)
        assert np.array_equal(eng.assign_grid(x0, y0, H, W), ref), (case, H, W, k)
        eng.accumulate()
        assert np.array_equal(eng.labels(), ref), (case, H, W, k)
        good = np.isfinite(w).ravel()
        rm = O.np_moments(pts[good], w.ravel()[good], ref.ravel()[good], k)
        np.testing.assert_allclose(eng.moments()[:, :4], rm[:, :4], rtol=1e-12, atol=1e-12, err_msg=str((case, H, W, k)))


def test_zero_weight_bin_updates_every_other_node(eng):
    """A bin with pixels but zero total weight (lloyd.pyx:66-74 gives NaN for it): every OTHER node
    must still be rewritten by that iteration, also when the table spans several CTAs of
    k_update_nodes (k > 256), and the error is reported once the iteration is complete."""
    N, k = 128, 700
    rng = np.random.default_rng(12)
    w = rng.uniform(0.5, 1.5, (N, N))
    w[:24, :24] = 0.0                                   # zero-density corner
    seeds = np.column_stack((rng.uniform(0, N - 1, k), rng.uniform(0, N - 1, k)))
    eng.set_image(density=w)
    eng.set_generators(seeds)
    eng.step(1)
    lab = O.c_assign_grid(N, N, seeds)
    x, y = np.meshgrid(np.arange(N, dtype=float), np.arange(N, dtype=float))
    rm = O.np_moments(np.column_stack((x.ravel(), y.ravel())), w.ravel(), lab.ravel(), k)
    with np.errstate(invalid="ignore", divide="ignore"):
        want, _ = O.np_update_nodes(rm, seeds)
    got = eng.generators()
    bad = ~np.isfinite(want).all(axis=1)
    assert bad.sum() > 0 and np.array_equal(~np.isfinite(got).all(axis=1), bad)
    np.testing.assert_allclose(got[~bad], want[~bad], rtol=0, atol=1e-9)
    assert eng.status() == (False, True) and eng.stats()[2] == 1
    eng.step(1)                                         # error flag up: further steps are no-ops
    assert eng.stats()[2] == 1


def test_long_lists_block_lists_and_their_overflow():
    """Quadrant lists of 65 .. 128 candidates (bins of 45 .. 110 px) take the 16x16 block lists; a block
    list that outgrows its capacity sends the quadrant through the exact per-pixel scan of the whole
    list.  That overflow needs > 64 generators around one 16x16 block, so it is forced here with a
    second emulator build whose block lists hold 16 entries (TESS_BCAP=16): labels bit-exact and moments
    to 1e-12 against the oracle for both builds.  Runs in subprocesses (one library per process)."""
    import os, subprocess, sys, textwrap
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    small = os.path.join(here, "emu", "_build", "libtess_emu_bcap16.so")
    srcs = [os.path.join(root, "tess_b200", "csrc", f) for f in os.listdir(os.path.join(root, "tess_b200", "csrc"))]
    if not os.path.exists(small) or any(os.path.getmtime(small) < os.path.getmtime(f) for f in srcs):
        os.makedirs(os.path.dirname(small), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++20", "-O2", "-pthread", "-DTESS_EMU", "-DTESS_BCAP=16", "-include",
                               os.path.join(here, "emu", "cuda_emu.h"), "-x", "c++",
                               os.path.join(root, "tess_b200", "csrc", "tess_b200.cu"), "-shared", "-fPIC", "-o", small])
    prog = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        from emu.harness import EmuEngine
        from oracle import oracle as O
        eng = EmuEngine()
        rng = np.random.default_rng(11)
        cases = []
        for area, scaled in ((50, False), (70, True), (100, False)):
            H, W = int(rng.integers(70, 120)), int(rng.integers(70, 120))
            k = int(H * W / area)
            g = np.column_stack((rng.uniform(-3, W + 3, k), rng.uniform(-3, H + 3, k)))
            g[: k // 10] = np.round(g[: k // 10])                       # some exact ties
            cases.append((H, W, g, rng.uniform(0.8, 1.3, k) if scaled else None))
        # 72 generators inside one 16x16 block of a 64x64 tile
        cases.append((64, 64, np.vstack((rng.uniform(16.2, 31.8, (72, 2)), rng.uniform(0, 63, (30, 2)))), None))
        for H, W, g, scale in cases:
            w = rng.uniform(0.5, 1.5, (H, W))
            x, y = np.meshgrid(np.arange(W, dtype=float), np.arange(H, dtype=float))
            pts = np.column_stack((x.ravel(), y.ravel()))
            ref = O.c_assign_points(pts, g, scale=scale).reshape(H, W)
            eng.set_image(density=w)
            eng.set_generators(g, scale=scale)
            eng.accumulate()
            assert np.array_equal(eng.labels(), ref), (H, W, len(g))
            rm = O.np_moments(pts, w.ravel(), ref.ravel(), len(g))
            assert np.allclose(eng.moments()[:, :4], rm[:, :4], rtol=1e-12, atol=1e-12), (H, W, len(g))
        print("ok")
        """) % (root, here)
    for lib in (None, small):
        env = dict(os.environ)
        env.pop("TESS_EMU_LIB", None)
        if lib:
            env["TESS_EMU_LIB"] = lib
        r = subprocess.run([sys.executable, "-c", prog], env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (lib, r.stdout[-500:], r.stderr[-1500:])
