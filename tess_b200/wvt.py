"""
tess_b200.wvt -- Lloyd/WVT iteration of a signal + noise map to a target S/N (SURVEY.md section 8f
rank 1): the loop `scripts/demo_iso_sn_voronoi.py:28-40` runs in two host steps (accretor seeds ->
`CVTessellation.from_image`), closed on the device with the per-bin S/N the reference defines in
`EqualSNAccretor.bin_sn` (tess/pixel_accretion.py:501-517): S/N_bin = sum(signal) / sqrt(sum(noise^2)).

Every iteration is the same fused kernel as plain Lloyd in S/N mode (6 moments per bin: count, sum w,
sum w x, sum w y, sum S, sum N^2).  With `scale_update` the node update also sets the WVT scale length
of every bin, s_j = sqrt(area_j / (S/N)_j) (Diehl & Statler 2006 eq. 4, as in vorbin), which the next
assignment uses in the scaled metric d^2 / s_j^2 -- all on the device, no host round trip.  The host
only looks at the S/N statistics every `check_every` iterations (one read-back of the k x 6 moment
table) to decide whether to stop.  The reference has no counterpart of the scale update; with
`scale_update=False` and `uniform_weight=False` the iteration is exactly `lloyd` with w = (S/N)^2.
"""
import numpy as np

from . import _capi
from .engine import Engine


def bin_sn(moments):
    """Per-bin S/N from the (k, 6) moment table; NaN for empty bins."""
    m = np.asarray(moments)
    with np.errstate(invalid="ignore", divide="ignore"):
        sn = m[:, 4] / np.sqrt(m[:, 5])
    return np.where(m[:, 0] > 0, sn, np.nan)


def sn_scatter(sn, target=None):
    """Fractional rms scatter of the bin S/N about `target` (about their own mean if None)."""
    s = np.asarray(sn)
    s = s[np.isfinite(s)]
    if s.size == 0:
        return np.nan
    t = float(np.mean(s)) if target is None else float(target)
    return float(np.sqrt(np.mean((s / t - 1.0) ** 2)))


def suggest_k(k, median_sn, target_sn):
    """Number of generators that would bring the median bin S/N to `target_sn`: the S/N of a bin grows
    like the square root of its area (sum S / sqrt(sum N^2) over ~equal pixels), i.e. like k^-1/2."""
    if not (np.isfinite(median_sn) and median_sn > 0 and target_sn and target_sn > 0):
        return None
    return max(1, int(round(k * (float(median_sn) / float(target_sn)) ** 2)))


def wvt_binning(signal, noise, generators, target_sn=None, max_iters=300, check_every=4, tol=1e-3,
                scale_update=True, uniform_weight=True, scale=None, engine=None, device=None, verbose=False,
                sn_tol=0.05, scatter_goal=None):
    """Iterate generators on a signal/noise map until the tessellation stops changing (the Lloyd
    latch, tess/lloyd.pyx:85-86), the S/N scatter about `target_sn` stops improving by more than `tol`
    (relative) between two checks, or `max_iters` + 2 iterations have run (the reference's budget,
    lloyd.pyx:87-90).

    `target_sn` is acted on: the loop also stops as soon as the target is MET -- the median bin S/N
    within `sn_tol` (relative) of it and, if `scatter_goal` is given, the fractional scatter about it at
    most that -- and when it cannot be met with this many generators the result says so:
    `target_met` False and `k_suggested`, the number of generators that would put the median on target
    (the reference's demo gets its k from the accretor run at the target, scripts/demo_iso_sn_voronoi.py:29).

    Returns a dict: nodes (k,2), scales (k,) or None, labels (H,W) int64, moments (k,6), bin_sn (k,),
    sn_scatter, median_sn, target_met (None without a target), k_suggested, n_iters, converged (latch
    closed), history [(iteration, scatter, median S/N)]."""
    eng = engine if engine is not None else Engine(device)
    eng.set_option(_capi.OPT_WVT_SCALE_UPDATE, 1 if scale_update else 0)
    eng.set_option(_capi.OPT_UNIFORM_WEIGHT, 1 if uniform_weight else 0)
    try:
        eng.set_image(signal=signal, noise=noise)
        eng.set_generators(generators, scale=scale)
        budget = max(int(max_iters) + 2, 1)
        done, conv, history, prev = 0, False, [], None
        while done < budget and not conv:
            n = min(int(check_every), budget - done)
            eng.step(n)
            done = eng.stats()[2]
            conv, bad = eng.status()
            if bad:
                raise _capi.TessNonFinite(-4, "a bin with pixels has zero total weight after iteration %d" % done)
            sn_now = bin_sn(eng._to_host(eng.moments()))
            sc = sn_scatter(sn_now, target_sn)
            med = float(np.nanmedian(sn_now)) if np.isfinite(sn_now).any() else np.nan
            history.append((done, sc, med))
            if verbose:
                print("WVT %03d  S/N scatter %.4f  median %.3f" % (done, sc, med))
            if target_sn is not None and np.isfinite(med) and abs(med / float(target_sn) - 1.0) <= sn_tol and \
                    (scatter_goal is None or sc <= scatter_goal):
                break                                  # on target
            if prev is not None and np.isfinite(sc) and abs(prev - sc) <= tol * max(prev, 1e-300):
                break
            prev = sc
        mom = eng._to_host(eng.moments())
        sn_fin = bin_sn(mom)
        med = float(np.nanmedian(sn_fin)) if np.isfinite(sn_fin).any() else np.nan
        met = None
        if target_sn is not None:
            met = bool(np.isfinite(med) and abs(med / float(target_sn) - 1.0) <= sn_tol and
                       (scatter_goal is None or sn_scatter(sn_fin, target_sn) <= scatter_goal))
        return dict(median_sn=med, target_met=met,
                    k_suggested=None if target_sn is None else suggest_k(int(mom.shape[0]), med, target_sn),
                    nodes=eng._to_host(eng.generators()),
                    scales=eng._to_host(eng.scales()) if (scale_update or scale is not None) else None,
                    labels=eng._to_host(eng.labels()).astype(np.int64), moments=mom, bin_sn=bin_sn(mom),
                    sn_scatter=sn_scatter(bin_sn(mom), target_sn), n_iters=int(done), converged=bool(conv),
                    history=history)
    finally:
        eng.set_option(_capi.OPT_WVT_SCALE_UPDATE, 0)
        eng.set_option(_capi.OPT_UNIFORM_WEIGHT, 0)
