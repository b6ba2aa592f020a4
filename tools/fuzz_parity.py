"""tools/fuzz_parity.py [seconds] [seed] -- randomised parity of the image path on the GPU: odd shapes,
non-integer window origins, generators inside / outside / on pixel centres / duplicated, bins from
thousands of pixels down to one, scaled metric on and off, all three kernel modes.  Labels must equal
the exhaustive O(n k) GPU scan bit for bit (that scan is itself pinned to the oracle by the test-suite);
moments must equal NumPy sums over those labels to 1e-12.  Prints one line per failure; exit code 1
if any."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tess_b200.engine import Engine
from tess_b200 import _capi

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
eng = Engine(0)
t0, n_case, n_fail = time.time(), 0, 0
while time.time() - t0 < budget:
    n_case += 1
    H, W = int(rng.integers(1, 700)), int(rng.integers(1, 700))
    x0, y0 = (float(rng.uniform(-50, 50)), float(rng.uniform(-50, 50))) if rng.random() < 0.5 else (0.0, 0.0)
    area = float(np.exp(rng.uniform(np.log(0.7), np.log(20000.0))))
    k = int(min(max(1, H * W / area), 300000))
    spread = rng.choice([1.0, 1.0, 1.3, 0.3])
    cx, cy = x0 + W / 2.0, y0 + H / 2.0
    if rng.random() < 0.5:
        g = np.column_stack((rng.uniform(cx - spread * W / 2, cx + spread * W / 2, k), rng.uniform(cy - spread * H / 2, cy + spread * H / 2, k)))
    else:
        g = np.column_stack((rng.normal(cx, spread * W / 6 + 0.5, k), rng.normal(cy, spread * H / 6 + 0.5, k)))
    if rng.random() < 0.3:
        m = max(1, k // 10)
        g[:m] = np.round(g[:m] - [x0, y0]) + [x0, y0]        # exactly on pixel centres (ties)
    if k > 3 and rng.random() < 0.3:
        g[k // 2] = g[1]                                      # exact duplicate
    scale = rng.uniform(0.6, 1.7, k) if rng.random() < 0.35 else None
    mode = int(rng.integers(0, 4))
    try:
        eng.set_option(_capi.OPT_UNIFORM_WEIGHT, 0)
        if rng.random() < 0.3:      # a rank's window inside a larger generator domain (sharded use)
            eng.set_domain(x0 - rng.uniform(0, 300), y0 - rng.uniform(0, 3000), x0 + W + rng.uniform(0, 300), y0 + H + rng.uniform(0, 3000))
        else:
            eng.set_domain(1.0, 1.0, 0.0, 0.0)      # none
        if mode == 3:               # point set (C2-like): lloyd on arbitrary xy
            n = int(rng.integers(1, 200000))
            pts = np.column_stack((rng.normal(cx, W / 4 + 1, n), rng.normal(cy, H / 5 + 1, n)))
            pw = rng.uniform(0.1, 1.0, n)
            eng.set_points(pts, pw)
            eng.set_generators(g, scale=scale)
            eng.accumulate()
            lab = eng.labels().cpu().numpy().ravel()
            ref = eng.assign_points(pts, exhaustive=True).cpu().numpy()
            ok = np.array_equal(lab, ref) and np.array_equal(eng.assign_points(pts).cpu().numpy(), ref)
            why = "point labels"
            if ok:
                want = np.column_stack((np.bincount(ref, minlength=k), np.bincount(ref, weights=pw, minlength=k),
                                        np.bincount(ref, weights=pw * pts[:, 0], minlength=k), np.bincount(ref, weights=pw * pts[:, 1], minlength=k)))
                got = eng.moments().cpu().numpy()[:, :4]
                ok = np.allclose(got, want, rtol=1e-11, atol=1e-9)
                why = "point moments (max abs diff %.3e)" % np.abs(got - want).max()
            if not ok:
                n_fail += 1
                print("FAIL case %d: %s n=%d k=%d scale=%s" % (n_case, why, n, k, scale is not None), flush=True)
            continue
        if mode == 2:
            sig = rng.uniform(0.5, 3.0, (H, W)); noise = rng.uniform(0.5, 1.5, (H, W))
            if rng.random() < 0.5: sig[rng.integers(0, H), rng.integers(0, W)] = np.nan
            eng.set_image(signal=sig, noise=noise, x0=x0, y0=y0)
            w = (sig / noise) ** 2
        else:
            w = rng.uniform(0.5, 1.5, (H, W))
            if rng.random() < 0.5: w[rng.integers(0, H), rng.integers(0, W)] = np.inf
            eng.set_image(density=w, x0=x0, y0=y0)
        eng.set_generators(g, scale=scale)
        ref = eng.assign_grid(x0, y0, H, W, exhaustive=True).cpu().numpy()
        if mode == 0:
            lab = eng.assign_grid(x0, y0, H, W).cpu().numpy()
            ok = np.array_equal(lab, ref)
            why = "segmap labels"
        else:
            eng.accumulate()
            lab = eng.labels().cpu().numpy()
            ok = np.array_equal(lab, ref)
            why = "labels"
            if ok:
                good = np.isfinite(w) if mode == 1 else (np.isfinite(w) & np.isfinite(sig) & np.isfinite(noise))
                x, y = np.meshgrid(x0 + np.arange(W, dtype=float), y0 + np.arange(H, dtype=float))
                l = ref[good]; ww = w[good]
                want = np.column_stack((np.bincount(l, minlength=k), np.bincount(l, weights=ww, minlength=k),
                                        np.bincount(l, weights=ww * x[good], minlength=k), np.bincount(l, weights=ww * y[good], minlength=k)))
                got = eng.moments().cpu().numpy()[:, :4]
                ok = np.allclose(got, want, rtol=1e-11, atol=1e-9)
                why = "moments (max abs diff %.3e)" % np.abs(got - want).max()
        if not ok:
            n_fail += 1
            bad = int((lab != ref).sum()) if lab.shape == ref.shape else -1
            print("FAIL case %d: %s  H=%d W=%d x0=%.3f y0=%.3f k=%d area=%.1f scale=%s mode=%d mismatches=%d stats=%s"
                  % (n_case, why, H, W, x0, y0, k, area, scale is not None, mode, bad, eng.list_stats()), flush=True)
    except Exception as e:
        n_fail += 1
        print("ERROR case %d: %r  H=%d W=%d k=%d mode=%d" % (n_case, e, H, W, k, mode), flush=True)
print("%d cases, %d failures in %.1f s" % (n_case, n_fail, time.time() - t0))
sys.exit(1 if n_fail else 0)
