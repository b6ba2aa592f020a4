"""
tess_b200.accretion -- the K2/K3 instances around the accretors.

The accretion loops themselves (tess/pixel_accretion.py:94-206, tess/point_accretion.pyx:32-68) are
sequential seed generators and stay on the CPU (out of scope).  What they hand to
the Lloyd path afterwards is device work and lives here, with the reference's semantics:

  segmap_centroids   PixelAccretor.centroids            tess/pixel_accretion.py:212-240
  segmap_bin_sn      EqualSNAccretor.bin_sn             tess/pixel_accretion.py:501-517
  cleanup_reassign   EqualSNAccretor.cleanup            tess/pixel_accretion.py:485-499
  point_nodes        PointAccretor.nodes                tess/point_accretion.pyx:70-84

Per-label sums run in libtess_b200.so (tess_label_sums), nearest-centroid re-assignment in
tess_assign_grid; there is no CPU fallback.  NumPy in -> NumPy out, torch CUDA in -> torch CUDA out
is NOT offered here: results are k-sized or image-sized host arrays, as in the reference.
"""
import numpy as np

from .engine import Engine


def _engine(engine, device):
    return engine if engine is not None else Engine(device)


def _label_moments(eng, seg, weights, n_bins):
    """count, sum w, sum w*row, sum w*col per label 0..n_bins-1 in one device pass with implicit pixel
    coordinates (tess_label_moments); pixels with non-finite weight are counted but carry no weight
    (pixel_accretion.py:229-232)."""
    lab = np.ascontiguousarray(seg, dtype=np.int32)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    m = eng._to_host(eng.label_moments(lab, w, n_bins))
    return m[:, 0], m[:, 1], m[:, 2], m[:, 3]


def segmap_centroids(segmap, weightmap=None, engine=None, device=None):
    """Weighted centroid of every bin of a segmentation image, as ``(x, y)`` rows.

    Mirrors `PixelAccretor.centroids` (tess/pixel_accretion.py:212-240) including its quirks: bins are
    ``range(segmap.max())`` (the highest-numbered bin is left out, :219,225), empty bins are dropped
    (:227-228,239-240), pixels with non-finite weight do not contribute (:230-233), and a bin whose
    finite weights sum to zero raises ZeroDivisionError as `np.average` does."""
    seg = np.asarray(segmap)
    n_bins = int(seg.max())
    if n_bins <= 0:
        return np.empty((0, 2))
    eng = _engine(engine, device)
    cnt, sw, swr, swc = _label_moments(eng, seg.astype(np.int64), weightmap, n_bins)
    have = cnt > 0
    if np.any(have & (sw == 0.0)):
        raise ZeroDivisionError("Weights sum to zero, can't be normalized")
    with np.errstate(invalid="ignore", divide="ignore"):
        cen = np.column_stack((swc / sw, swr / sw))            # swapped to (x, y) order (:235-238)
    cen[~have] = np.nan
    return cen[np.isfinite(cen[:, 0])]


def segmap_bin_sn(segmap, image, noise, engine=None, device=None):
    """S/N of each non-empty bin, ``sum(image) / sqrt(sum(noise**2))`` over its pixels
    (`EqualSNAccretor.bin_sn`, tess/pixel_accretion.py:501-517; bins ``range(segmap.max())``)."""
    seg = np.asarray(segmap)
    n_bins = int(seg.max())
    if n_bins <= 0:
        return np.empty((0,))
    eng = _engine(engine, device)
    lab = eng._to_dev(np.ascontiguousarray(seg, dtype=np.int32), "int32")
    cnt = eng._to_host(eng.label_sums(lab, None, n_bins))
    s = eng._to_host(eng.label_sums(lab, eng._to_dev(np.asarray(image, dtype=np.float64)), n_bins))
    n2 = eng._to_host(eng.label_sums(lab, eng._to_dev(np.asarray(noise, dtype=np.float64) ** 2.0), n_bins))
    have = cnt > 0
    with np.errstate(invalid="ignore", divide="ignore"):
        return s[have] / np.sqrt(n2[have])


def cleanup_reassign(segmap, bin_centroids, valid_bins, engine=None, device=None):
    """Merge failed bins into their neighbours: every pixel of a failed bin takes the label of the
    good bin whose centroid is nearest (`EqualSNAccretor.cleanup`, tess/pixel_accretion.py:485-499).

    `bin_centroids` is the accretor's ``_current_bin_centroids``, one ``(row, col)`` per bin -- the
    reference queries its kd-tree in (row, col) space (:497-498); `valid_bins` the boolean per bin.
    Returns a new segmentation image (the reference edits it in place).  Exact distance ties go to
    the lowest good-bin index (the reference's choice depends on kd-tree order)."""
    seg = np.array(segmap, copy=True)
    valid = np.asarray(valid_bins, dtype=bool)
    cen = np.asarray(bin_centroids, dtype=np.float64)
    good = np.where(valid)[0]
    failed = np.where(~valid)[0]
    if failed.size == 0 or good.size == 0:
        return seg
    H, W = seg.shape
    eng = _engine(engine, device)
    # generators as (x, y) = (col, row): the same squared distance, the two squares added in the
    # other order
    eng.set_generators(np.ascontiguousarray(cen[good][:, ::-1]))
    nearest = eng._to_host(eng.assign_grid(0.0, 0.0, H, W)).astype(np.int64)
    bad = np.isin(seg, failed)
    seg[bad] = good[nearest[bad]]
    return seg


def point_nodes(xy, w, bin_nums, n_bins, engine=None, device=None):
    """Mass-weighted centroid of every point bin (`PointAccretor.nodes`,
    tess/point_accretion.pyx:70-84): `bin_nums` are 1-based, bins without mass give NaN as the
    reference's 0/0 does."""
    xy = np.asarray(xy, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    eng = _engine(engine, device)
    lab = eng._to_dev((np.asarray(bin_nums, dtype=np.int64) - 1).astype(np.int32), "int32")
    m = eng._to_host(eng.label_sums(lab, eng._to_dev(w), int(n_bins)))
    mx = eng._to_host(eng.label_sums(lab, eng._to_dev(xy[:, 0] * w), int(n_bins)))
    my = eng._to_host(eng.label_sums(lab, eng._to_dev(xy[:, 1] * w), int(n_bins)))
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.column_stack((mx / m, my / m))
