import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import gauss_seeds, device_signal_rows
from tess_b200.engine import Engine
from tess_b200.sharded import shard_rows
N, k = 32768, 1000000
fr, fn = (int(x) for x in sys.argv[1].split("/"))
r0, r1 = shard_rows(N, fn)[fr]
dev = torch.device("cuda", 0)
dens = device_signal_rows(N, r0, r1, torch, dev) ** 2
eng = Engine(0)
eng.set_image(density=dens, y0=float(r0))
eng.set_domain(0.0, 0.0, float(N), float(N))
eng.set_generators(gauss_seeds(N, k))
for it in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.step(1); b.record(); torch.cuda.synchronize()
    print(it, a.elapsed_time(b), eng.list_stats(), eng.stats())
