"""
tess_b200.lloyd -- drop-in for the reference's `tess.lloyd.lloyd` (tess/lloyd.pyx:10-90).

    node_xy, idx, converged = lloyd(xy, w, node_xy, max_iters)

Same positional signature, same return triple, same stop rule (at most max_iters + 2 loop bodies,
`converged` True only at an exact fixed point, lloyd.pyx:85-90), same error behaviour for wrong
dtypes (the reference's typed memoryviews raise ValueError for anything but float64) and for a
zero-weight bin (ValueError, as cKDTree raises on the NaN node one iteration later).  Inputs are
not modified.  NumPy in -> NumPy out (idx int64, like the reference's `long`); torch CUDA tensors
in -> torch CUDA tensors out, with no host copy of the point set.

The work is done by libtess_b200.so on the GPU (tess_b200/engine.py); there is no CPU fallback.
Differences from the reference, all documented in DESIGN.md:
  * the per-iteration `print "CVT Delta %03d %.2e"` (lloyd.pyx:82) is off unless verbose=True;
  * exact distance ties go to the lowest generator index (cKDTree's choice is tree-order
    dependent);
  * per-bin sums are formed in a different order, so nodes agree to ~1e-13 px, not bitwise.
"""
import numpy as np

from .engine import Engine

_ENGINES = {}


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _check_f64(name, a, ndim):
    dt = str(a.dtype).replace("torch.", "")
    if dt != "float64":
        # the reference's `double[:, :]` memoryview raises ValueError("Buffer dtype mismatch ...")
        raise ValueError("Buffer dtype mismatch, expected 'double' but got '%s' (argument %s)" % (dt, name))
    if a.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d) (argument %s)"
                         % (ndim, a.ndim, name))


def shared_engine(device=None):
    """One cached Engine per device: repeated lloyd() calls reuse its workspace."""
    import torch
    idx = None if device is None else torch.device(device).index
    if idx is None:                      # 'cuda' / an index-less tensor device: the current device, not GPU 0
        idx = torch.cuda.current_device()
    eng = _ENGINES.get(idx)
    if eng is None or eng.ctx is None:
        eng = _ENGINES[idx] = Engine(idx)
    return eng


def run_lloyd(eng, max_iters, verbose=False):
    """The reference loop on a prepared engine.  Returns (converged, n_iters)."""
    if not verbose:
        return eng.run(max_iters)
    # step by step so that the reference's per-iteration line (lloyd.pyx:82) can be printed
    budget = max(int(max_iters) + 2, 1)
    done = 0
    while done < budget:
        eng.run(-1)                      # exactly one loop body
        delta, _, it = eng.stats()
        print("CVT Delta %03d %.2e" % (done, delta))
        done += 1
        if delta == 0.0:
            return True, done
    return False, done


def lloyd(xy, w, node_xy, max_iters, verbose=False, device=None, engine=None):
    """Lloyd's algorithm (reference tess/lloyd.pyx:10-35 for the contract)."""
    as_torch = _is_torch(xy)
    if not as_torch:
        xy, w, node_xy = np.asarray(xy), np.asarray(w), np.asarray(node_xy)
    _check_f64("xy", xy, 2)
    _check_f64("w", w, 1)
    _check_f64("node_xy", node_xy, 2)
    n, k = int(xy.shape[0]), int(node_xy.shape[0])
    if k == 0:
        raise ValueError("Invalid shape in axis 0: 0.")            # the reference's memoryview error
    if int(xy.shape[1]) != 2 or int(node_xy.shape[1]) != 2:
        raise ValueError("x must consist of vectors of length 2 but has shape %s"
                         % (tuple(xy.shape) if int(xy.shape[1]) != 2 else tuple(node_xy.shape),))
    # the reference indexes w[i] for i < n (lloyd.pyx:63): a longer w is read partially, a shorter
    # one is an IndexError from its bounds check
    if int(w.shape[0]) < n:
        raise IndexError("Out of bounds on buffer access (axis 0)")
    if int(w.shape[0]) > n:
        w = w[:n]
    if n == 0:
        # no points: every bin is empty, nodes go to (0, 0) (lloyd.pyx:71-74) and the second
        # iteration finds delta == 0 (lloyd.pyx:85-86); nothing to compute
        if as_torch:
            return xy.new_zeros((k, 2)), xy.new_zeros((0,), dtype=xy.long().dtype), True
        return np.zeros((k, 2)), np.zeros((0,), dtype=np.int64), True
    eng = engine if engine is not None else shared_engine(device if device is not None else
                                                          (xy.device if as_torch else None))
    eng.set_points(xy, w)
    eng.set_generators(node_xy)
    converged, _ = run_lloyd(eng, int(max_iters), verbose)
    nodes = eng.generators()
    idx = eng.labels()
    if as_torch:
        return nodes, idx.long(), converged
    return eng._to_host(nodes), eng._to_host(idx).astype(np.int64), converged
