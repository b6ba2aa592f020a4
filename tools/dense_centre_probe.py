"""tools/dense_centre_probe.py -- a C3-like map whose centre is unbinned (one generator per pixel in a
D x D square): ms per accumulate against the plain C3 generator table."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tess_b200.engine import Engine
from bench import gauss_seeds
N = 4096
eng = Engine(0)
dens = torch.rand((N, N), dtype=torch.float64, device="cuda") + 0.5
eng.set_image(density=dens)
base = gauss_seeds(N, 25000)
for D in (0, 64, 256, 512, 1024):
    node = base
    if D:
        c0 = (N - D) // 2
        yy, xx = np.mgrid[c0:c0 + D, c0:c0 + D]
        node = np.vstack((base, np.column_stack((xx.ravel(), yy.ravel())).astype(np.float64)))
    eng.set_generators(node)
    eng.accumulate(); eng.accumulate()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        eng.accumulate()
    b.record(); torch.cuda.synchronize()
    print(json.dumps(dict(D=D, k=int(node.shape[0]), ms_accumulate=a.elapsed_time(b) / 5, list_stats=eng.list_stats())), flush=True)
