"""
Parity of the CUDA path (through the C ABI, via tess_b200.engine / lloyd / voronoi) with
 (i) the golden fixtures generated from the reference itself (tests/golden/make_golden.py),
 (ii) the CPU oracle on seeded inputs at sizes it finishes in seconds, and
 (iii) size-independent properties at BASELINE.json's full sizes.
Bars: labels bit-exact; per-bin moments <= 1e-12 relative; converged nodes <= 1e-6 px.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import golden, gaussian_image, sampled_seeds
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture()
def eng(engine_cls):
    e = engine_cls(0)
    yield e
    e.close()


def h(t):
    return t.cpu().numpy()


# ---------------------------------------------------------------- golden fixtures (reference outputs)
def test_img64_trajectory_matches_reference(eng):
    g = golden("lloyd_img64.npz")
    eng.set_image(density=g["density"])
    eng.set_generators(g["seeds"])
    T = int(g["n_iters"])
    for t in range(T):
        eng.step(1)
        assert np.array_equal(h(eng.labels()).ravel(), g["labels"][t].astype(np.int32)), t
        assert np.abs(h(eng.generators()) - g["nodes"][t]).max() < 1e-9, t
        delta = eng.stats()[0]
        assert abs(delta - float(g["deltas_txt"][t])) <= 6e-3 * max(delta, 1e-300) or (delta == 0.0 and t == T - 1)
    assert eng.status() == (True, False)
    eng.step(3)
    assert eng.stats()[2] == T
    for max_iters, want in ((300, (True, T)), (3, (False, 5)), (0, (False, 2))):
        eng.set_generators(g["seeds"])
        assert eng.run(max_iters) == want


def test_config_c1_demo256(engine_cls):
    """C1: scripts/demo_iso_sn_voronoi.py inputs, EqualSNAccretor seeds, Lloyd to convergence."""
    from tess_b200.voronoi import CVTessellation
    g = golden("lloyd_demo256.npz")
    dens = gaussian_image(256) ** 2.0
    cv = CVTessellation.from_image(dens, g["seeds"])
    assert np.abs(cv.nodes - g["nodes"]).max() < 1e-6
    assert np.array_equal(cv.membership, g["membership"]) and cv.membership.dtype == np.int64
    assert np.array_equal(cv.segmap, g["segmap"])
    assert cv._engine.stats()[2] == int(g["n_iters"])


def test_config_c2_shape_points_fixture():
    from tess_b200.lloyd import lloyd
    from tess_b200.voronoi import CVTessellation
    g = golden("lloyd_points.npz")
    n, i, c = lloyd(g["xy"], g["w"], g["seeds"], int(g["max_iters"]))
    assert not c and np.array_equal(i, g["membership"]) and i.dtype == np.int64
    assert np.abs(n - g["nodes"]).max() < 1e-6
    cv = CVTessellation(g["xy"], g["w"], node_xy=g["seeds"], max_iters=int(g["max_iters"]))
    np.testing.assert_allclose(cv.node_weights, g["node_weights"], rtol=1e-12)
    assert np.array_equal(cv.membership, g["membership"])


def test_voronoi_api_fixture():
    from tess_b200.voronoi import VoronoiTessellation
    g = golden("segmap.npz")
    vt = VoronoiTessellation(g["gens"])
    vt.set_pixel_grid(tuple(g["xlim"]), tuple(g["ylim"]))
    assert np.array_equal(vt.segmap, g["segmap"]) and vt.segmap.dtype == np.int64
    assert np.array_equal(vt.render_voronoi_field(g["values"]), g["field"])
    assert np.array_equal(vt.partition_points(g["points"]), g["partition"])
    np.testing.assert_allclose(vt.sum_cell_point_mass(g["points"], mass=g["mass"]), g["cell_mass"], rtol=1e-13)
    assert np.array_equal(vt.compute_cell_areas(), np.bincount(g["segmap"].ravel()))
    dens = vt.cell_point_density(g["points"], mass=g["mass"])
    np.testing.assert_allclose(dens, g["cell_mass"] / np.bincount(g["segmap"].ravel()), rtol=1e-13)
    # flagmap > 0 pixels are left out of the areas (docstring semantics of tess/voronoi.py:116-152)
    import frow_checks
    frow_checks.check_flagmap(vt, g)


def test_from_image_nan_fixture():
    from tess_b200.voronoi import CVTessellation
    g = golden("from_image_nan.npz")
    cv = CVTessellation.from_image(g["density"], g["seeds"])
    assert np.abs(cv.nodes - g["nodes"]).max() < 1e-6
    assert np.array_equal(cv.membership, g["membership"])
    np.testing.assert_allclose(cv.node_weights, g["node_weights"], rtol=1e-12)
    assert np.array_equal(cv.segmap, g["segmap"])


def test_torch_in_torch_out():
    import torch
    from tess_b200.lloyd import lloyd
    g = golden("lloyd_points.npz")
    xy = torch.from_numpy(g["xy"]).cuda(); w = torch.from_numpy(g["w"]).cuda(); s = torch.from_numpy(g["seeds"]).cuda()
    n, i, c = lloyd(xy, w, s, 5)
    rn, ri, rc, _, _, _ = O.c_lloyd(g["xy"], g["w"], g["seeds"], 5)
    assert n.is_cuda and i.is_cuda and i.dtype == torch.int64 and c == rc
    assert np.array_equal(h(i), ri) and np.abs(h(n) - rn).max() < 1e-9
    assert torch.equal(s, torch.from_numpy(g["seeds"]).cuda())          # inputs are not modified
    with pytest.raises(ValueError):
        lloyd(xy.float(), w, s, 3)


# ---------------------------------------------------------------- oracle on seeded inputs
def test_1024_map_teacher_forced_iterations(eng):
    """1024^2 x 1563 tie-free: labels bit-exact, moments 1e-12, next nodes 1e-9."""
    N, k = 1024, 1563
    w = gaussian_image(N) ** 2.0
    seeds = sampled_seeds(w, k, seed=0)
    x, y = np.meshgrid(np.arange(N, dtype=float), np.arange(N, dtype=float))
    xy = np.column_stack((x.ravel(), y.ravel()))
    eng.set_image(density=w)
    node = seeds
    for it in range(3):
        eng.set_generators(node)          # teacher forcing: the oracle's node table goes in
        eng.accumulate()
        lab = h(eng.labels())
        ref = O.c_assign_grid(N, N, node)
        assert np.array_equal(lab, ref), it
        mom = h(eng.moments())
        rm = O.np_moments(xy, w.ravel(), ref.ravel(), k)
        np.testing.assert_allclose(mom[:, :4], rm[:, :4], rtol=1e-12, atol=0)
        eng.update()
        node, _ = O.np_update_nodes(rm, node)
        assert np.abs(h(eng.generators()) - node).max() < 1e-9


def test_sn_mode_moments_and_mask(eng):
    sig = gaussian_image(300) + 1.0
    noise = np.sqrt(sig + 1.0)
    bad = np.random.default_rng(2).choice(300 * 300, 90, replace=False)
    sig.ravel()[bad[:45]] = np.nan
    noise.ravel()[bad[45:]] = np.inf
    w = (sig / noise) ** 2
    seeds = sampled_seeds(np.nan_to_num(w, posinf=0.0), 120, seed=8)
    eng.set_image(signal=sig, noise=noise)
    eng.set_generators(seeds)
    eng.accumulate()
    lab = h(eng.labels())
    assert np.array_equal(lab, O.c_assign_grid(300, 300, seeds))
    good = np.isfinite(w) & np.isfinite(sig) & np.isfinite(noise)
    x, y = np.meshgrid(np.arange(300.0), np.arange(300.0))
    xy = np.column_stack((x.ravel(), y.ravel()))
    idx = np.where(good.ravel(), lab.ravel(), -1).astype(np.int64)
    ref = O.np_moments(xy, np.nan_to_num(w.ravel(), posinf=0.0), idx, 120, signal=np.nan_to_num(sig.ravel()),
                       noise=np.nan_to_num(noise.ravel(), posinf=0.0))
    mom = h(eng.moments())
    np.testing.assert_allclose(mom, ref, rtol=1e-12, atol=0)
    np.testing.assert_allclose(mom[:, 4] / np.sqrt(mom[:, 5]), O.bin_sn(ref), rtol=1e-12)


def test_wvt_scales_ties_far_generators_odd_shapes(eng):
    rng = np.random.default_rng(6)
    node = rng.uniform(0, 400, (900, 2))
    scale = rng.uniform(0.5, 3.0, 900)
    eng.set_generators(node, scale)
    lab = h(eng.assign_grid(-3.0, 2.0, 301, 417))
    assert np.array_equal(lab, O.c_assign_grid(301, 417, node, x0=-3.0, y0=2.0, scale=scale))
    pts = rng.uniform(-50, 450, (20000, 2))
    assert np.array_equal(h(eng.assign_points(pts)), O.c_assign_points(pts, node, scale))
    # exact ties -> lowest index
    gx, gy = np.meshgrid(np.arange(0.5, 200, 4.0), np.arange(0.5, 200, 4.0))
    tn = np.column_stack((gx.ravel(), gy.ravel()))
    tn = tn[np.random.default_rng(1).permutation(tn.shape[0])]
    eng.set_generators(tn)
    assert np.array_equal(h(eng.assign_grid(0, 0, 200, 200)), O.c_assign_grid(200, 200, tn))
    # all generators in one corner; windows away from them; odd widths (scalar load/store path)
    cn = rng.uniform(0, 6, (25, 2))
    eng.set_generators(cn)
    for (x0, y0, H, W) in ((0, 0, 530, 267), (50.5, -20.25, 65, 129), (-3000, 4000, 3, 5)):
        assert np.array_equal(h(eng.assign_grid(x0, y0, H, W)), O.c_assign_grid(H, W, cn, x0=x0, y0=y0))
    # more generators than a tile list can hold -> exhaustive-scan fallback inside the kernel
    dense = rng.uniform(0, 64, (3000, 2))
    eng.set_generators(dense)
    assert np.array_equal(h(eng.assign_grid(0, 0, 64, 64)), O.c_assign_grid(64, 64, dense))
    # every point its own generator (CVTessellation default, voronoi.py:298-299)
    from tess_b200.lloyd import lloyd
    pts = rng.uniform(0, 30, (500, 2))
    n, i, c = lloyd(pts, np.ones(500), pts.copy(), 10)
    assert c and np.array_equal(i, np.arange(500)) and np.abs(n - pts).max() < 1e-12


def test_empty_bins_and_zero_weight_bin(eng):
    from tess_b200.lloyd import lloyd
    xy = np.array([[5.0, 5.0], [6.0, 5.0], [5.0, 6.0], [7.0, 7.5]])
    n, i, c = lloyd(xy, np.ones(4), np.array([[5.2, 5.1], [100.0, 100.0]]), 0)
    rn, ri, rc, _, _, _ = O.c_lloyd(xy, np.ones(4), np.array([[5.2, 5.1], [100.0, 100.0]]), 0)
    assert np.array_equal(i, ri) and np.allclose(n, rn, atol=1e-12) and c == rc      # empty -> (0,0)
    xy = np.array([[0.0, 0.0], [1.0, 0.0], [10.0, 0.0], [11.0, 0.0]])
    with pytest.raises(ValueError):
        lloyd(xy, np.array([1.0, 1.0, 1.0, -1.0]), np.array([[0.4, 0.1], [10.4, 0.1]]), 10)
    with pytest.raises(ValueError):
        eng.set_generators(np.array([[np.nan, 1.0]]))


def test_fake_sharding_equals_single_engine(engine_cls):
    """Two contexts on one GPU, each with half of the rows; their moment vectors are summed as the
    all-reduce would (tess_b200/sharded.py) -- labels and nodes must equal the unsharded run."""
    N, k = 512, 400
    w = gaussian_image(N) ** 2.0
    seeds = sampled_seeds(w, k, seed=3)
    one, a, b = engine_cls(0), engine_cls(0), engine_cls(0)
    one.set_image(density=w); one.set_generators(seeds)
    a.set_image(density=w[:256], y0=0.0); a.set_domain(0.0, 0.0, N, N); a.set_generators(seeds)
    b.set_image(density=w[256:], y0=256.0); b.set_domain(0.0, 0.0, N, N); b.set_generators(seeds)
    for it in range(4):
        one.step(1)
        a.accumulate(); b.accumulate()
        va, _ = a.moments_view(); vb, _ = b.moments_view()
        tot = va + vb
        va.copy_(tot); vb.copy_(tot)
        a.update(); b.update()
        lab = np.vstack((h(a.labels()), h(b.labels())))
        assert np.array_equal(lab, h(one.labels())), it
        np.testing.assert_allclose(h(a.generators()), h(one.generators()), rtol=0, atol=1e-9)
        assert np.array_equal(h(a.generators()), h(b.generators()))
    for e in (one, a, b):
        e.close()


# ---------------------------------------------------------------- full-size properties
def _device_gaussian(N, torch):
    s = 50.0 * N / 256.0
    ax = torch.arange(N, dtype=torch.float64, device="cuda")
    gx = torch.exp(-((ax - N / 2.0) ** 2) / (2.0 * s * s))
    return (10.0 * gx[:, None] * gx[None, :]) ** 2


def _gauss_seeds(N, k, seed):
    rng = np.random.default_rng(seed)
    s = 50.0 * N / 256.0 / np.sqrt(2.0)
    out = np.empty((0, 2))
    while out.shape[0] < k:
        c = rng.normal(N / 2.0, s, (2 * k, 2))
        c = c[(c[:, 0] > 0) & (c[:, 0] < N - 1) & (c[:, 1] > 0) & (c[:, 1] < N - 1)]
        out = np.vstack((out, c))
    return out[:k]


def test_config_c3_full_size_properties(eng):
    """8192^2 x 1e5 (headline): conservation, Lloyd-path == segmap-path, tiled == exhaustive on a
    window, determinism of labels, label-change latch."""
    import torch
    N, k = 8192, 100000
    dens = _device_gaussian(N, torch)
    seeds = _gauss_seeds(N, k, 0)
    eng.set_image(density=dens)
    eng.set_generators(seeds)
    eng.accumulate()
    lab = eng.labels()
    mom = eng.moments()
    assert float(mom[:, 0].sum().item()) == float(N) * N                       # every pixel counted once
    tot = dens.sum().item()
    assert abs(mom[:, 1].sum().item() - tot) <= 1e-11 * tot                    # weight conserved
    ax = torch.arange(N, dtype=torch.float64, device="cuda")
    wx = (dens * ax[None, :]).sum().item(); wy = (dens * ax[:, None]).sum().item()
    assert abs(mom[:, 2].sum().item() - wx) <= 1e-11 * wx and abs(mom[:, 3].sum().item() - wy) <= 1e-11 * wy
    assert int(lab.min().item()) >= 0 and int(lab.max().item()) < k
    # per-bin count == bincount of the label map (a checksum of checksums)
    assert torch.equal(torch.bincount(lab.reshape(-1).long(), minlength=k).double(), mom[:, 0])
    # the segmap path (tess_assign_grid) and the Lloyd path agree everywhere
    assert torch.equal(eng.assign_grid(0.0, 0.0, N, N), lab)
    # tiled == exhaustive O(n k) on a window through the dense centre and on a far corner
    for (x0, y0, H, W) in ((3584, 3840, 512, 768), (0, 0, 256, 384)):
        ex = eng.assign_grid(float(x0), float(y0), H, W, exhaustive=True)
        assert torch.equal(ex, lab[y0:y0 + H, x0:x0 + W])
    # same generators again: labels identical, and the moment sums agree to 1e-12
    eng.set_generators(seeds)
    eng.accumulate()
    assert torch.equal(eng.labels(), lab)
    assert torch.allclose(eng.moments(), mom, rtol=1e-12, atol=0)
    # oracle on a 64-row strip
    strip = O.c_assign_grid(8, 512, seeds, x0=3840.0, y0=4096.0)
    assert np.array_equal(h(lab[4096:4104, 3840:4352]), strip)
    # a few free-running iterations: delta decreases overall, node table stays finite
    eng.step(3)
    d, _, it = eng.stats()
    assert it == 3 and np.isfinite(d) and d > 0


def test_config_c2_full_size_points(eng):
    """1e6 points x 1e4 generators (C2): tiled == exhaustive, oracle on a subset, conservation."""
    import torch
    rng = np.random.default_rng(1)
    n, k = 1000000, 10000
    xy = np.column_stack((np.clip(rng.normal(250, 100, n), 0, 500), np.clip(rng.normal(250, 50, n), 0, 500)))
    w = rng.uniform(0.1, 1.0, n)
    seeds = xy[rng.choice(n, k, replace=False)] + rng.uniform(-1e-3, 1e-3, (k, 2))
    eng.set_points(xy, w)
    eng.set_generators(seeds)
    eng.accumulate()
    lab = eng.labels()
    assert torch.equal(lab, eng.assign_points(xy, exhaustive=True))
    sub = rng.choice(n, 20000, replace=False)
    assert np.array_equal(h(lab)[sub], O.c_assign_points(xy[sub], seeds))
    mom = h(eng.moments())
    assert mom[:, 0].sum() == n and abs(mom[:, 1].sum() - w.sum()) <= 1e-11 * w.sum()
    rm = O.np_moments(xy, w, h(lab).astype(np.int64), k)
    np.testing.assert_allclose(mom[:, :4], rm[:, :4], rtol=1e-12, atol=0)
    conv, nit = eng.run(48)
    assert nit == 50 or conv


def test_config_c5_segmap_large(eng):
    """Segmentation-map rasterisation at a large size (C5 shape scaled to fit the test budget:
    16384^2 x 2.5e5 generators): tiled == exhaustive on windows, all labels valid."""
    import torch
    N, k = 16384, 250000
    seeds = _gauss_seeds(N, k, 5)
    eng.set_generators(seeds)
    seg = eng.assign_grid(0.0, 0.0, N, N)
    assert int(seg.min().item()) >= 0 and int(seg.max().item()) < k
    for (x0, y0, H, W) in ((8000, 8100, 256, 512), (16000, 100, 128, 384)):
        ex = eng.assign_grid(float(x0), float(y0), H, W, exhaustive=True)
        assert torch.equal(ex, seg[y0:y0 + H, x0:x0 + W])
    # idempotence: the same call returns the same map
    assert torch.equal(eng.assign_grid(0.0, 0.0, N, N), seg)


def _bincount_moments(torch, lab, w, N, y0=0):
    """Independent device-side per-bin sums (torch.bincount) of count, w, w x, w y for a [H, N] map."""
    l = lab.reshape(-1).long()
    ax = torch.arange(N, dtype=torch.float64, device=lab.device)
    ay = torch.arange(y0, y0 + lab.shape[0], dtype=torch.float64, device=lab.device)
    out = [torch.bincount(l), torch.bincount(l, weights=w.reshape(-1)),
           torch.bincount(l, weights=(w * ax[None, :]).reshape(-1)),
           torch.bincount(l, weights=(w * ay[:, None]).reshape(-1))]
    return out


def test_config_c3_sn_mode_full_size(eng):
    """C3 in signal + noise mode (6 moments, 20 B/px) at full size: labels equal the CVT-mode labels,
    all six moments against torch.bincount over the whole map, per-bin S/N, NaN/Inf mask."""
    import torch
    N, k = 8192, 100000
    sig = torch.sqrt(_device_gaussian(N, torch)) + 1.0
    noise = torch.sqrt(sig + 1.0)
    bad = torch.from_numpy(np.random.default_rng(2).choice(N * N, N * N // 1000, replace=False)).cuda()
    sig.view(-1)[bad[::2]] = float("nan")
    noise.view(-1)[bad[1::2]] = float("inf")
    seeds = _gauss_seeds(N, k, 0)
    eng.set_image(signal=sig, noise=noise)
    eng.set_generators(seeds)
    eng.accumulate()
    lab, mom = eng.labels(), eng.moments()
    assert torch.equal(lab, eng.assign_grid(0.0, 0.0, N, N))
    w = (sig / noise) ** 2
    good = torch.isfinite(w) & torch.isfinite(sig) & torch.isfinite(noise)
    assert float(mom[:, 0].sum().item()) == float(good.sum().item())
    zero = torch.zeros((), dtype=torch.float64, device="cuda")
    wz, sz, nz = torch.where(good, w, zero), torch.where(good, sig, zero), torch.where(good, noise, zero)
    ref = _bincount_moments(torch, lab, wz, N)
    ref[0] = torch.bincount(lab.reshape(-1).long(), weights=good.reshape(-1).double())
    ref += [torch.bincount(lab.reshape(-1).long(), weights=sz.reshape(-1)),
            torch.bincount(lab.reshape(-1).long(), weights=(nz * nz).reshape(-1))]
    for c, r in enumerate(ref):
        r = torch.nn.functional.pad(r.double(), (0, k - r.numel()))
        assert torch.allclose(mom[:, c], r, rtol=1e-11, atol=0), c
    x0, y0, H, W = 3900, 4000, 96, 128
    sub = O.c_assign_grid(H, W, seeds, x0=float(x0), y0=float(y0))
    assert np.array_equal(h(lab[y0:y0 + H, x0:x0 + W]), sub)


def _full_size_checks(torch, eng, N, k, lab, mom, dens, seeds, windows, strips):
    assert int(lab.min().item()) >= 0 and int(lab.max().item()) < k
    if mom is not None:
        assert float(mom[:, 0].sum().item()) == float(N) * N
        ref = _bincount_moments(torch, lab, dens, N)
        assert torch.equal(torch.nn.functional.pad(ref[0], (0, k - ref[0].numel())).double(), mom[:, 0])
        for c in (1, 2, 3):
            r = torch.nn.functional.pad(ref[c], (0, k - ref[c].numel()))
            assert torch.allclose(mom[:, c], r, rtol=1e-11, atol=0), c
        del ref
    for (x0, y0, H, W) in windows:
        ex = eng.assign_grid(float(x0), float(y0), H, W, exhaustive=True)
        assert torch.equal(ex, lab[y0:y0 + H, x0:x0 + W]), (x0, y0)
    for (x0, y0, H, W) in strips:
        assert np.array_equal(h(lab[y0:y0 + H, x0:x0 + W]), O.c_assign_grid(H, W, seeds, x0=float(x0), y0=float(y0)))


def test_config_c4_full_size_one_gpu(eng):
    """C4 (32768^2 px x 1e6 generators, CVT mode) on ONE GPU: conservation, per-bin sums against
    torch.bincount over the whole map, Lloyd path == segmap path, tiled == exhaustive on windows
    through the dense centre and a far corner, oracle strips."""
    import torch
    N, k = 32768, 1000000
    dens = _device_gaussian(N, torch)
    seeds = _gauss_seeds(N, k, 0)
    eng.set_image(density=dens)
    eng.set_generators(seeds)
    eng.accumulate()
    lab, mom = eng.labels(), eng.moments()
    _full_size_checks(torch, eng, N, k, lab, mom, dens, seeds,
                      windows=((16300, 16350, 160, 256), (0, 0, 128, 256), (32512, 20000, 128, 256)),
                      strips=((16000, 16384, 4, 256), (30000, 3000, 4, 256)))
    seg = eng.assign_grid(0.0, 0.0, N, N)
    assert torch.equal(seg, lab)
    del seg
    eng.step(2)
    d, _, it = eng.stats()
    assert it == 2 and np.isfinite(d) and d > 0


def test_config_c5_full_size(eng):
    """C5 at BASELINE's size: segmentation map of 32768^2 px x 1e6 generators."""
    import torch
    N, k = 32768, 1000000
    seeds = _gauss_seeds(N, k, 5)
    eng.set_generators(seeds)
    seg = eng.assign_grid(0.0, 0.0, N, N)
    _full_size_checks(torch, eng, N, k, seg, None, None, seeds,
                      windows=((16384, 16300, 160, 256), (32000, 200, 128, 256)),
                      strips=((16100, 16500, 4, 256), (100, 32700, 4, 256)))
    cnt = torch.bincount(seg.reshape(-1).long(), minlength=k)
    assert int(cnt.sum().item()) == N * N
    assert torch.equal(eng.assign_grid(0.0, 0.0, N, N), seg)             # idempotence


# ---------------------------------------------------------------- WVT extension, options, diagnostics
def test_wvt_scaled_lloyd_iterations_vs_oracle(eng):
    """Scale-weighted metric in the Lloyd loop (no reference counterpart: parity is against the
    restatement, oracle/oracle.py, which degenerates bit-exactly to the reference at scale == 1)."""
    N, k = 384, 260
    sig = gaussian_image(N) + 1.0
    noise = np.sqrt(sig + 1.0)
    w = (sig / noise) ** 2
    seeds = sampled_seeds(w, k, seed=4)
    scale = np.random.default_rng(9).uniform(0.7, 1.8, k)
    x, y = np.meshgrid(np.arange(N, dtype=float), np.arange(N, dtype=float))
    xy = np.column_stack((x.ravel(), y.ravel()))
    eng.set_image(signal=sig, noise=noise)
    node = seeds
    for it in range(3):
        eng.set_generators(node, scale)
        eng.accumulate()
        lab = h(eng.labels())
        ref = O.c_assign_grid(N, N, node, scale=scale)
        assert np.array_equal(lab, ref), it
        rm = O.np_moments(xy, w.ravel(), ref.ravel(), k, signal=sig.ravel(), noise=noise.ravel())
        np.testing.assert_allclose(h(eng.moments()), rm, rtol=1e-12, atol=0)
        eng.update()
        node, _ = O.np_update_nodes(rm, node)
        assert np.abs(h(eng.generators()) - node).max() < 1e-9
        assert np.array_equal(h(eng.scales()), scale)
    # scale == 1 is bit-identical to no scales at all
    eng.set_generators(seeds, np.ones(k)); eng.accumulate(); a = h(eng.labels())
    eng.set_generators(seeds); eng.accumulate()
    assert np.array_equal(a, h(eng.labels()))


def test_options_uniform_weight_and_scale_update(eng):
    from tess_b200 import _capi
    N, k = 256, 120
    sig = gaussian_image(N) + 1.0
    noise = np.sqrt(sig + 1.0)
    seeds = sampled_seeds((sig / noise) ** 2, k, seed=6)
    x, y = np.meshgrid(np.arange(N, dtype=float), np.arange(N, dtype=float))
    xy = np.column_stack((x.ravel(), y.ravel()))
    eng.set_option(_capi.OPT_UNIFORM_WEIGHT, 1)         # geometric centroids: w = 1
    eng.set_option(_capi.OPT_WVT_SCALE_UPDATE, 1)       # scale_j = sqrt(count_j / (S/N)_j) after each update
    eng.set_image(signal=sig, noise=noise)
    eng.set_generators(seeds)
    eng.step(1)
    lab = O.c_assign_grid(N, N, seeds)
    assert np.array_equal(h(eng.labels()), lab)
    rm = O.np_moments(xy, np.ones(N * N), lab.ravel(), k, signal=sig.ravel(), noise=noise.ravel())
    np.testing.assert_allclose(h(eng.moments()), rm, rtol=1e-12, atol=0)
    sn = O.bin_sn(rm)
    np.testing.assert_allclose(h(eng.scales()), np.sqrt(rm[:, 0] / sn), rtol=1e-13)
    # the next iteration uses the updated scales
    node, _ = O.np_update_nodes(rm, seeds)
    eng.step(1)
    assert np.array_equal(h(eng.labels()), O.c_assign_grid(N, N, node, scale=np.sqrt(rm[:, 0] / sn)))
    eng.set_option(_capi.OPT_UNIFORM_WEIGHT, 0)
    eng.set_option(_capi.OPT_WVT_SCALE_UPDATE, 0)


def test_list_diagnostics_and_dense_fallbacks(eng):
    """More generators than pixels per quadrant can hold on the fast path: the scan fallbacks are
    exact; tess_get_list_stats reports them."""
    rng = np.random.default_rng(3)
    N = 192
    dense = rng.uniform(0, N, (N * N // 20, 2))          # one generator per 20 px
    w = rng.uniform(0.5, 1.5, (N, N))
    x, y = np.meshgrid(np.arange(N, dtype=float), np.arange(N, dtype=float))
    xy = np.column_stack((x.ravel(), y.ravel()))
    eng.set_image(density=w)
    eng.set_generators(dense)
    eng.accumulate()
    ref = O.c_assign_grid(N, N, dense)
    assert np.array_equal(h(eng.labels()), ref)
    np.testing.assert_allclose(h(eng.moments())[:, :4], O.np_moments(xy, w.ravel(), ref.ravel(), dense.shape[0])[:, :4],
                               rtol=1e-12, atol=0)
    st = eng.list_stats()
    assert st["ring_fallback"] + st["scan_list"] > 0 and st["arena_used"] <= st["arena_capacity"]


@pytest.mark.parametrize("area", [40.0, 10.0, 2.0, 1.0])
def test_small_bins_ring_fallback(eng, area):
    """Bins of a few pixels: tiles overflow the list builder and run the per-pixel ring search over a
    refined generator grid.  Labels bit-exact against the exhaustive GPU scan (itself checked against
    the oracle on a corner), moments against NumPy sums of the same labels."""
    rng = np.random.default_rng(int(area))
    N = 256
    k = int(N * N / area)
    node = rng.uniform(-0.5, N - 0.5, (k, 2))
    node[: k // 50] = np.round(node[: k // 50])           # some generators exactly on pixel centres
    node[5] = node[4]                                     # an exact duplicate: lowest index wins
    w = rng.uniform(0.5, 1.5, (N, N))
    eng.set_image(density=w)
    eng.set_generators(node)
    eng.accumulate()
    lab = h(eng.labels())
    ref = h(eng.assign_grid(0.0, 0.0, N, N, exhaustive=True))
    assert np.array_equal(lab, ref)
    assert np.array_equal(ref[:24, :24], O.c_assign_grid(N, N, node)[:24, :24])
    assert not np.any(lab == 5)
    x, y = np.meshgrid(np.arange(N, dtype=float), np.arange(N, dtype=float))
    rm = O.np_moments(np.column_stack((x.ravel(), y.ravel())), w.ravel(), ref.ravel(), k)
    np.testing.assert_allclose(h(eng.moments())[:, :4], rm[:, :4], rtol=1e-12, atol=0)
    st = eng.list_stats()
    assert st["ring_fallback"] > 0
    # the segmap query takes the same path
    assert np.array_equal(h(eng.assign_grid(0.0, 0.0, N, N)), ref)


def test_dense_centre_sparse_outskirts(eng):
    """Unbinned pixels in the centre (one generator per pixel), large bins outside: the grid is refined
    from the fullest tile, not the mean occupancy; labels stay bit-exact everywhere."""
    rng = np.random.default_rng(11)
    N = 512
    c0, c1 = 200, 312
    yy, xx = np.mgrid[c0:c1, c0:c1]
    centre = np.column_stack((xx.ravel(), yy.ravel())).astype(np.float64)
    outer = rng.uniform(-0.5, N - 0.5, (300, 2))
    outer = outer[(np.abs(outer - 255.5) > 70).any(axis=1)]
    node = np.vstack((outer, centre))
    w = rng.uniform(0.5, 1.5, (N, N))
    eng.set_image(density=w)
    eng.set_generators(node)
    eng.accumulate()
    lab = h(eng.labels())
    ref = h(eng.assign_grid(0.0, 0.0, N, N, exhaustive=True))
    assert np.array_equal(lab, ref)
    assert np.array_equal(lab[c0:c1, c0:c1].ravel(), outer.shape[0] + np.arange(centre.shape[0]))
    x, y = np.meshgrid(np.arange(N, dtype=float), np.arange(N, dtype=float))
    rm = O.np_moments(np.column_stack((x.ravel(), y.ravel())), w.ravel(), ref.ravel(), node.shape[0])
    np.testing.assert_allclose(h(eng.moments())[:, :4], rm[:, :4], rtol=1e-12, atol=0)


# ---------------------------------------------------------------- SURVEY.md section 8f rows
def test_accretion_helpers_gpu(engine_cls):
    import frow_checks
    frow_checks.check_accretion_helpers(lambda: engine_cls(0))
    frow_checks.check_accretor_fixture(lambda: engine_cls(0))


def test_wvt_binning_to_target_sn_gpu(engine_cls):
    import frow_checks
    r = frow_checks.check_wvt_binning(lambda: engine_cls(0), N=256, k=120)
    assert r["labels"].shape == (256, 256)


def test_external_moment_buffer_and_peer_allreduce_world1(eng, engine_cls):
    """tess_set_moment_buffer + tess_peer_allreduce on one GPU: caller-owned moment vector gives the
    same moments, and the (world = 1) two-shot kernel is the identity on it.  The multi-rank paths
    are exercised by tools/peer_probe.py and bench.py --gpus N on the multi-GPU boxes."""
    import ctypes
    import torch
    N, k = 192, 60
    w = gaussian_image(N) + 0.5
    seeds = sampled_seeds(w, k, seed=2)
    eng.set_image(density=w)
    eng.set_generators(seeds)
    eng.step(2)
    want_m, want_g = h(eng.moments()), h(eng.generators())
    e2 = engine_cls(0)
    buf = torch.full((6 * k + 2,), 7.0, dtype=torch.float64, device="cuda")
    e2.set_image(density=w)
    e2.set_generators(seeds)
    e2.set_moment_buffer(buf, buf.numel())
    ptr, n, m = e2.api.moments_ptr(e2.ctx)
    assert ptr == buf.data_ptr() and m == 4 and n == 4 * k + 2
    ptrs = (ctypes.c_void_p * 1)(buf.data_ptr())
    for _ in range(2):
        e2.accumulate()
        before = buf[:n].clone()
        e2.peer_allreduce(1, 0, ptrs, 0, n)
        torch.cuda.synchronize()
        assert torch.equal(before, buf[:n])
        e2.update()
    np.testing.assert_allclose(h(e2.moments()), want_m, rtol=1e-12, atol=0)
    np.testing.assert_allclose(h(e2.generators()), want_g, rtol=0, atol=1e-9)
    with pytest.raises(Exception):
        e2.set_moment_buffer(torch.zeros(8, dtype=torch.float64, device="cuda"), 8)
        e2.accumulate()
    e2.set_moment_buffer(None, 0)
    e2.accumulate()
    e2.close()


def test_sparse_exchange_and_owner_computed_update_world1(engine_cls):
    """tess_peer_allreduce_sparse on one GPU (world = 1: the rank owns the whole vector), with rows of
    totals and with TESS_OPT_OWNER_UPDATE -- the owner forms the new nodes (lloyd.pyx:66-74) and
    tess_lloyd_update takes them from the node area of the caller's buffer: generators equal to
    the single-context iteration (1e-9 px; labels bit-identical), empty bins to (0, 0) included.  k > 512 so that the slice spans
    several 16 KB chunks; the multi-rank runs are checked by bench.py --gpus N (parity_check)."""
    import ctypes
    import torch
    from tess_b200 import _capi
    N, k = 384, 1500
    w = gaussian_image(N) + 0.5
    seeds = sampled_seeds(w, k, seed=4)
    seeds[7] = [-500.0, -500.0]                 # never nearest: an empty bin
    e1 = engine_cls(0)
    e1.set_image(density=w)
    e1.set_generators(seeds)
    e1.step(3)
    want_g, want_m = e1.generators().clone(), e1.moments().clone()
    assert torch.equal(want_g[7], torch.zeros(2, dtype=torch.float64, device="cuda")) or float(want_m[7, 0]) > 0
    for owner in (0, 1):
        e2 = engine_cls(0)
        buf = torch.zeros((6 * k + 2 + 34,), dtype=torch.float64, device="cuda")
        e2.set_image(density=w)
        e2.set_generators(seeds)
        e2.set_moment_buffer(buf, buf.numel())
        e2.set_option(_capi.OPT_OWNER_UPDATE, owner)
        e2.peer_set_multicast(0)                  # (no multicast mapping on one GPU: peer-to-peer copies)
        ptr, n, m = e2.api.moments_ptr(e2.ctx)
        ptrs = (ctypes.c_void_p * 1)(buf.data_ptr())
        for _ in range(3):
            e2.accumulate()
            e2.peer_allreduce_sparse(1, 0, ptrs, n)
            e2.update()
        torch.cuda.synchronize()
        # (two contexts sum in different orders: last-bit differences in the moments, none in the labels)
        np.testing.assert_allclose(h(e2.generators()), h(want_g), rtol=0, atol=1e-9, err_msg=str(owner))
        assert torch.equal(e2.labels(), e1.labels()), owner
        np.testing.assert_allclose(h(e2.moments()), h(want_m), rtol=1e-12, atol=0, err_msg=str(owner))   # (world = 1: the owner's copy holds every total)
        assert e2.stats()[2] == 3
        e2.set_moment_buffer(None, 0)
        e2.close()
    e1.close()


def test_randomised_parity_short():
    """tools/fuzz_parity.py for a few seconds: random shapes, origins, densities (thousands of px per
    bin down to one), ties, duplicates, scales and modes against the exhaustive scan."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_parity.py"), "8", "7"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " 0 failures" in r.stdout


def test_nonfinite_points_raise_like_ckdtree(eng):
    """lloyd.pyx:55 / voronoi.py:171-172: cKDTree.query raises ValueError("'x' must be finite")."""
    from tess_b200.lloyd import lloyd
    from tess_b200.voronoi import VoronoiTessellation
    g = np.array([[1., 2.], [3., 4.]])
    xy = np.random.default_rng(0).uniform(0, 5, (1000, 2))
    xy[417, 1] = np.nan
    with pytest.raises(ValueError):
        lloyd(xy, np.ones(1000), g, 10, engine=eng)
    vt = VoronoiTessellation(g, engine=eng)
    with pytest.raises(ValueError):
        vt.partition_points(np.array([[0.5, np.inf], [1.0, 1.0]]))
    assert np.array_equal(vt.partition_points(np.array([[0.5, 2.0], [3.1, 4.2]])), [0, 1])
    n, i, c = lloyd(np.empty((0, 2)), np.empty(0), g, 10, engine=eng)
    assert np.array_equal(n, np.zeros((2, 2))) and i.shape == (0,) and c is True
