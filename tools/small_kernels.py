"""tools/small_kernels.py -- per-kernel device times of one Lloyd iteration at k = 1e6 on a small
image (the O(k) kernels dominate): tuning aid for the replicated per-rank work of multi-GPU runs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tess_b200.engine import Engine
from bench import gauss_seeds
N, k = 32768, 1000000
eng = Engine(0)
dens = torch.rand((2048, N), dtype=torch.float64, device="cuda") + 0.5
eng.set_image(density=dens, y0=15360.0)
eng.set_domain(0.0, 0.0, float(N), float(N))
eng.set_generators(gauss_seeds(N, k))
for _ in range(3):
    eng.accumulate()
torch.cuda.synchronize()
a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
a.record(); eng.accumulate(); b.record(); eng.update(); c.record(); torch.cuda.synchronize()
print("accumulate %.3f ms  update %.3f ms" % (a.elapsed_time(b), b.elapsed_time(c)))
