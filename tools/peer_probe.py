"""tools/peer_probe.py -- torchrun --nproc-per-node N tools/peer_probe.py: is symmetric memory usable
here, does the own-kernel exchange give the same sums as NCCL, and what does each cost at the C4
payload (4e6 + 2 doubles)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from tess_b200.engine import Engine
from tess_b200.sharded import ShardedLloyd

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
k = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
eng = Engine(local)
rng = np.random.default_rng(0)
node = rng.uniform(0, 4096, (k, 2))
dens = torch.rand((256, 4096), dtype=torch.float64, device="cuda") + 0.5
for mode in ("peer", "peer_sparse_mc", "nccl"):
    os.environ["TESS_EXCHANGE"] = "nccl" if mode == "nccl" else "peer"
    os.environ["TESS_SPARSE_MC"] = "1" if mode == "peer_sparse_mc" else "0"
    try:
        drv = ShardedLloyd(eng)
        drv.set_image_rows(density=dens, row0=256 * rank, total_rows=256 * world)
        drv.set_generators(node)
    except Exception as e:
        if rank == 0:
            print(mode, "FAILED:", repr(e)[:300], flush=True)
        continue
    if rank == 0:
        print(mode, "->", drv.exchange, flush=True)
    mv = drv._moments()
    n = mv.numel()
    g = torch.Generator(device="cuda"); g.manual_seed(100 + rank)
    src = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    mv.copy_(src)
    ref = src.clone()
    dist.all_reduce(ref)
    torch.cuda.synchronize(); dist.barrier()
    drv.exchange_moments()
    torch.cuda.synchronize()
    err = float(((mv - ref).abs() / ref.abs().clamp_min(1e-300)).max())
    # bit-identical on all ranks?
    chk = mv.view(torch.int64).sum().reshape(1).clone()
    lst = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    same = all(int(x) == int(lst[0]) for x in lst)
    for _ in range(5):
        drv.exchange_moments()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        drv.exchange_moments()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / 20], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%s: n=%d  max rel err vs NCCL %.2e  identical on all ranks %s  %.1f us per exchange" % (mode, n, err, same, 1e3 * float(t)), flush=True)
    if drv._peer is not None:
        h = drv._peer["hdl"]
        a.record()
        for _ in range(20):
            h.barrier(channel=0); h.barrier(channel=1)
        b.record(); torch.cuda.synchronize()
        if rank == 0:
            print("  two barriers alone: %.1f us" % (1e3 * a.elapsed_time(b) / 20), flush=True)
    # and a real sharded step stays consistent
    drv.set_generators(node)
    drv.step(2)
    torch.cuda.synchronize()
    gsum = drv.generators().sum().reshape(1).clone()
    lst = [torch.zeros_like(gsum) for _ in range(world)]
    dist.all_gather(lst, gsum)
    if rank == 0:
        print("  after 2 steps: node checksum", float(lst[0]), "equal on all ranks:", all(float(x) == float(lst[0]) for x in lst), flush=True)
    eng.set_moment_buffer(None, 0)
eng.close()
dist.destroy_process_group()
