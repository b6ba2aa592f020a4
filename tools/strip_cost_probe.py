"""
tools/strip_cost_probe.py -- cost of one accumulate (list builder + image kernel) per strip of rows of
the C4 map on ONE GPU, with the generators relaxed by a few Lloyd iterations first: the data the
row-shard cost model of tess_b200/sharded.py::balanced_row_shards is fitted to.  Tuning aid.
    python tools/strip_cost_probe.py [strip_rows=512] [relax_iterations=8] > gpurun_out/strips.json
"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import gauss_seeds, device_signal_rows
from tess_b200.engine import Engine

N, k = 32768, 1000000
SR = int(sys.argv[1]) if len(sys.argv) > 1 else 512
RELAX = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
dens = device_signal_rows(N, 0, N, torch, dev) ** 2
eng = Engine(0)
eng.set_image(density=dens)
eng.set_generators(gauss_seeds(N, k))
eng.step(RELAX)
torch.cuda.synchronize()
gen = eng.generators().cpu().numpy().copy()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); eng.step(5); b.record(); torch.cuda.synchronize()
full_ms = a.elapsed_time(b) / 5
full_acc_ms = None
del eng
out = {"N": N, "k": k, "strip_rows": SR, "relax": RELAX, "full_iter_ms": full_ms, "full_accumulate_ms": full_acc_ms, "strips": []}
eng = Engine(0)
for r0 in range(0, N, SR):
    r1 = min(N, r0 + SR)
    eng.set_image(density=dens[r0:r1], y0=float(r0))
    eng.set_domain(0.0, 0.0, float(N), float(N))
    eng.set_generators(gen)
    # one accumulate right after set_generators: no previous labels, so no label comparison rides along
    # (the index build it includes is the same for every strip and ends up in the fit's constant)
    ts = []
    for rep in range(4):
        eng.set_generators(gen)
        torch.cuda.synchronize()
        a.record(); eng.accumulate(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    out["strips"].append([r0, r1, float(np.median(ts[1:]))])
np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "relaxed_generators_c4.npy"), gen.astype(np.float32))
print(json.dumps(out))
