"""
tools/regime_bench.py -- ms per Lloyd iteration of the image kernel as a function of bin size
(uniform random generators on a flat-ish map), to see which density regime of the composite
headline workload costs what.  Tuning aid, not part of the product or the test-suite.
"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tess_b200.engine import Engine
from tess_b200 import _capi

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = Engine(0)
rng = np.random.default_rng(0)
dens = torch.rand((N, N), dtype=torch.float64, device="cuda") + 0.5
eng.set_image(density=dens)
out = []
AREAS = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [80, 160, 320, 700, 1500, 4000, 12000, 40000, 150000]
for area in AREAS:
    k = max(4, int(N * N / area))
    seeds = rng.uniform(0, N - 1, (k, 2))
    eng.set_generators(seeds)
    eng.step(3)
    torch.cuda.synchronize()
    eng.set_option(_capi.OPT_KERNEL_TIMING, 1)
    eng.kernel_timing()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.step(10); b.record(); torch.cuda.synchronize()
    kms, kn = eng.kernel_timing()
    eng.set_option(_capi.OPT_KERNEL_TIMING, 0)
    ms = a.elapsed_time(b) / 10
    out.append(dict(area=area, k=k, ms_iter=ms, ms_main=kms / kn, ns_per_px_main=1e6 * kms / kn / (N * N),
                    gpix_s=N * N / (kms / kn * 1e-3) / 1e9))
    print(json.dumps(out[-1]))
