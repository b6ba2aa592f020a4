"""
tools/other_configs.py -- timings of the non-headline configurations of BASELINE.json on one GPU:
C2 (point-set CVT: 1e6 points, 1e4 generators, 50 Lloyd iterations) and C5 (segmentation-map
rasterisation only: 32768^2 px x 1e6 generators).  Prints one JSON object per configuration.
"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tess_b200.engine import Engine
from bench import gauss_seeds

eng = Engine(0)

# ---- C2: examples/starfield_test.py scaled
rng = np.random.default_rng(1)
n, k = 1000000, 10000
xy = np.column_stack((np.clip(rng.normal(250, 100, n), 0, 500), np.clip(rng.normal(250, 50, n), 0, 500)))
w = rng.uniform(0.1, 1.0, n)
seeds = xy[rng.choice(n, k, replace=False)] + rng.uniform(-1e-3, 1e-3, (k, 2))
dxy, dw = torch.from_numpy(xy).cuda(), torch.from_numpy(w).cuda()
eng.set_points(dxy, dw)
eng.set_generators(seeds)
eng.step(3)
torch.cuda.synchronize()
eng.set_generators(seeds)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); conv, nit = eng.run(48); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
print(json.dumps({"config": "C2 point-set CVT 1e6 pts x 1e4 gens", "iterations": nit, "converged": conv,
                  "ms_total": ms, "ms_per_iter": ms / nit, "Mpt_per_s": n / (ms / nit * 1e-3) / 1e6,
                  "algorithmic_GBps": 28.0 * n / (ms / nit * 1e-3) / 1e9}))

# ---- C5: segmap only
N, k = 32768, 1000000
eng.set_generators(gauss_seeds(N, k))
out = eng.assign_grid(0.0, 0.0, N, N)
torch.cuda.synchronize()
del out
a.record(); out = eng.assign_grid(0.0, 0.0, N, N); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
print(json.dumps({"config": "C5 segmap 32768^2 px x 1e6 gens", "ms": ms, "Gpix_per_s": N * N / (ms * 1e-3) / 1e9,
                  "write_GBps": 4.0 * N * N / (ms * 1e-3) / 1e9, "frac_of_hbm_6509": 4.0 * N * N / (ms * 1e-3) / 1e9 / 6509.1}))
