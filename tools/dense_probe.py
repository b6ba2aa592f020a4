"""tools/dense_probe.py -- where the small-bin cliff is: ms per iteration and fallback counters versus
bin area on a 1024^2 map (bounded work even when every quadrant scans the whole table)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tess_b200.engine import Engine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
AREAS = [float(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [100, 60, 40, 30, 20, 10, 5, 2, 1]
eng = Engine(0)
rng = np.random.default_rng(0)
dens = torch.rand((N, N), dtype=torch.float64, device="cuda") + 0.5
eng.set_image(density=dens)
for area in AREAS:
    k = max(4, int(N * N / area))
    eng.set_generators(rng.uniform(0, N - 1, (k, 2)))
    eng.accumulate()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        eng.accumulate()
    b.record(); torch.cuda.synchronize()
    print(json.dumps(dict(area=area, k=k, ms_accumulate=a.elapsed_time(b) / 5, list_stats=eng.list_stats())), flush=True)
