"""
tools/shard_cost_probe.py -- accumulate time of each of the 8 row shards of the C4 map on ONE GPU
(generators relaxed by a few iterations), under the current cost model of balanced_row_shards and
under the previous one: checks the balance before an 8-GPU run is spent on it.  Tuning aid.
"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import gauss_seeds, device_signal_rows
from tess_b200.engine import Engine
from tess_b200.sharded import balanced_row_shards

N, k, W = 32768, 1000000, 8
dev = torch.device("cuda", 0)
dens = device_signal_rows(N, 0, N, torch, dev) ** 2
eng = Engine(0)
eng.set_image(density=dens)
seeds = gauss_seeds(N, k)
eng.set_generators(seeds)
eng.step(8)
torch.cuda.synchronize()
gen = eng.generators().cpu().numpy().copy()
del eng
eng = Engine(0)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, model in (("current", None), ("previous", (2.56, 39.1, 0.40))):
    cuts = balanced_row_shards(seeds, N, N, W, model=model)      # (the bench cuts on the seeds)
    ts = []
    for r0, r1 in cuts:
        eng.set_image(density=dens[r0:r1], y0=float(r0))
        eng.set_domain(0.0, 0.0, float(N), float(N))
        t = []
        for rep in range(4):
            eng.set_generators(gen)
            torch.cuda.synchronize()
            a.record(); eng.accumulate(); b.record(); torch.cuda.synchronize()
            t.append(a.elapsed_time(b))
        ts.append(float(np.median(t[1:])))
    print(json.dumps({"model": name, "cuts": cuts, "accumulate_ms": [round(x, 4) for x in ts],
                      "max_over_mean": max(ts) / np.mean(ts)}))
