#!/usr/bin/env python
"""
bench.py -- throughput of the Lloyd / WVT iteration (BASELINE.json: "WVT Lloyd ms/iter and Gpix/s,
8192^2 px x 1e5 bins; 1/2/4/8 B200 vs HBM roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c4|small] [--mode cvt|sn]

A "step" is one Lloyd iteration (grid index rebuild + assignment + per-bin sums + node update) over
the whole synthetic map.  N = 1: configs[2], 8192^2 px x 1e5 generators (headline).  N > 1 (under
torchrun, one rank per GPU): configs[3], 32768^2 px x 1e6 generators, pixel rows sharded across
ranks, one NCCL all-reduce of the per-bin moment vector per iteration (strong scaling: the map is
fixed); rank 0 additionally times the same map on one GPU so the speed-up is apples to apples.
Synthetic data (no network): the demo's Gaussian S/N map scaled to N, generators
drawn from the same Gaussian (p ~ w) with continuous coordinates (tie-free).

Prints ONE JSON line (rank 0).  `value` = Gpix/s with inputs resident in HBM; `e2e` = the same
through the public API with pinned HOST buffers (H2D of the map + generators and D2H of labels +
nodes inside the timed region); `roofline` = algorithmic bytes of the dominant kernel / its device
time measured live with CUDA events, against MEASURED_PEAKS.json; `cpu_baseline` = the compiled
reference (oracle/_ref, tess/lloyd.pyx unmodified) timed on a bounded slab of the same workload.
`--impl reference` times only that CPU reference.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "lloyd_wvt_iteration_throughput"
UNIT = "Gpix/s"
WORKLOADS = {
    "c3": dict(N=8192, k=100000, name="synthetic 8192^2 px S/N map, 1e5 bins (configs[2], headline)"),
    "c4": dict(N=32768, k=1000000, name="synthetic 32768^2 px S/N map, 1e6 bins, row-sharded (configs[3])"),
    "small": dict(N=2048, k=6250, name="synthetic 2048^2 px S/N map, 6250 bins (debug)"),
}
BYTES_PER_PX = {"cvt": 12.0, "sn": 20.0}     # 8 (density) + 4 (label) | 8 + 8 + 4


# ------------------------------------------------------------------------------------ synthetic data
def sigma(N):
    return 50.0 * N / 256.0


def gauss_seeds(N, k, seed=0):
    """k generators drawn with density ~ w = (S/N)^2 (a Gaussian of std sigma/sqrt2), inside the map."""
    rng = np.random.default_rng(seed)
    s = sigma(N) / np.sqrt(2.0)
    out = np.empty((0, 2))
    while out.shape[0] < k:
        c = rng.normal(N / 2.0, s, (2 * k, 2))
        c = c[(c[:, 0] > 0) & (c[:, 0] < N - 1) & (c[:, 1] > 0) & (c[:, 1] < N - 1)]
        out = np.vstack((out, c))
    return np.ascontiguousarray(out[:k])


def device_signal_rows(N, r0, r1, torch, device):
    """Rows [r0, r1) of signal[y, x] = 10 exp(-((x-N/2)^2 + (y-N/2)^2) / (2 sigma^2)); noise == 1."""
    s = sigma(N)
    ax = torch.arange(N, dtype=torch.float64, device=device)
    gx = torch.exp(-((ax - N / 2.0) ** 2) / (2.0 * s * s))
    return 10.0 * gx[r0:r1, None] * gx[None, :]


def host_signal_rows(N, r0, r1):
    s = sigma(N)
    ax = np.arange(N, dtype=np.float64)
    gx = np.exp(-((ax - N / 2.0) ** 2) / (2.0 * s * s))
    return 10.0 * gx[r0:r1, None] * gx[None, :]


# ------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    """Samples SM clock and throttle reasons of one GPU DURING the timed region: NVML in a thread
    (every ~2 ms; the timed region is tens of milliseconds), nvidia-smi -lms as a fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = None
        self.p = None
        self.th = None
        self.stop_flag = False
        self.active = False      # samples are kept only while the timed region runs
        self.sm, self.mx, self.reasons, self.power = [], [], set(), []

    def _nvml_loop(self):
        import pynvml as nv
        h = self.h
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            if not self.active:
                time.sleep(0.0005)
                continue
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for nm, b in bits.items():
                    if r & b:
                        self.reasons.add(nm)
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates in PCI order; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                idx = int(vis.split(",")[self.index])
            self.h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx = [float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))]
            self.th = threading.Thread(target=self._nvml_loop, daemon=True)
            self.th.start()
            return
        except Exception:
            self.th = None
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.th is not None:
            self.stop_flag = True
            self.th.join(1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                    "sm_max_mhz": max(self.mx) if self.mx else None, "samples": len(self.sm),
                    "power_w_max": max(self.power) if self.power else None, "reasons": sorted(self.reasons),
                    "source": "nvml"}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------ CPU reference
def reference_lloyd_callable():
    """(callable(xy, w, nodes, max_iters) -> ..., kind).  The compiled reference when present
    (oracle/_ref: tess/lloyd.pyx built unmodified), else the oracle's C port."""
    from oracle import oracle as O
    if O.ref_available():
        try:
            O.ref_module()
            return (lambda xy, w, nd, mi: O.ref_lloyd(xy, w, nd, mi)), "reference"
        except Exception:
            pass
    return (lambda xy, w, nd, mi: O.c_lloyd(xy, w, nd, mi)), "port"


def cpu_slab(N, k, rows, seed=0):
    """A centred slab of `rows` image rows of the workload as the reference's point set
    (tess/voronoi.py:281-287) with ALL k generators."""
    r0 = max(0, N // 2 - rows // 2)
    r1 = min(N, r0 + rows)
    sig = host_signal_rows(N, r0, r1)
    w = (sig / 1.0) ** 2
    x, y = np.meshgrid(np.arange(N, dtype=float), np.arange(r0, r1, dtype=float))
    xy = np.column_stack((x.ravel(), y.ravel()))
    return xy, np.ascontiguousarray(w.ravel()), gauss_seeds(N, k, seed), (r0, r1)


def time_reference(N, k, rows, iters):
    """Seconds per reference iteration on the slab (1 thread: lloyd.pyx:55 passes no `workers`)."""
    fn, kind = reference_lloyd_callable()
    xy, w, seeds, rr = cpu_slab(N, k, rows)
    t0 = time.perf_counter()
    fn(xy, w, seeds, iters - 2)                   # max_iters = t - 2 runs exactly t loop bodies
    dt = (time.perf_counter() - t0) / iters
    return dt, xy.shape[0], kind, rr


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, k = wl["N"], wl["k"]
    total = args.steps + args.warmup
    # ~1.5 Mpix/s on one core: size the slab so that the whole run takes about two minutes
    px_budget = 1.5e6 * 120.0 / max(total, 1)
    rows = int(min(N, max(16, px_budget // N)))
    fn, kind = reference_lloyd_callable()
    xy, w, seeds, rr = cpu_slab(N, k, rows)
    for _ in range(args.warmup):
        fn(xy, w, seeds, -1)                      # max_iters = -1: exactly one loop body
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        fn(xy, w, seeds, -1)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    val = xy.shape[0] / (ms * 1e-3) / 1e9
    sample = "rows %d..%d of the %dx%d map (%d px) x all %d generators, 1 Lloyd iteration per step" % (
        rr[0], rr[1], N, N, xy.shape[0], k)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "mode": args.mode, "sample": sample, "host_cpus": os.cpu_count()},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=[None] + sorted(WORKLOADS))
    ap.add_argument("--mode", default="cvt", choices=["cvt", "sn"],
                    help="cvt: density map, 4 moments, 12 B/px (reference parity); sn: signal+noise maps, 6 moments, 20 B/px")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample clocks during the timed region")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-single", action="store_true", help="N>1: skip rank 0's one-GPU run of the same map (and the parity check against it)")
    ap.add_argument("--no-extras", action="store_true", help="N=1: skip the other-mode and the sustained measurements")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="N=1: length of the sustained run in seconds")
    ap.add_argument("--no-balance", dest="balance", action="store_false",
                    help="N>1: equal-height row shards instead of cost-balanced ones")
    ap.add_argument("--no-rebalance", action="store_true",
                    help="N>1: keep the cost-model row shards (no second cut from measured accumulate times)")
    ap.add_argument("--trace-steps", action="store_true",
                    help="diagnostic: after the timed region, time a second run of --steps steps one by one (stderr)")
    ap.add_argument("--fake-shard", default=None, metavar="R/N",
                    help="profiling aid on ONE GPU: process only the rows of rank R of N (no collective)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl_key = args.workload or ("c3" if max(world, args.gpus) == 1 else "c4")
    wl = WORKLOADS[wl_key]
    if args.impl == "reference":
        return run_reference_arm(args, wl)

    import torch
    import torch.distributed as dist
    from tess_b200.engine import Engine
    from tess_b200 import _capi
    from tess_b200.sharded import ShardedLloyd, shard_rows, balanced_row_shards

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    # stdout carries exactly ONE JSON line: park fd 1 on stderr while libraries (NCCL prints its
    # version banner to stdout) initialise and run, restore it for the final print
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    N, k, mode = wl["N"], wl["k"], args.mode
    seeds_h = gauss_seeds(N, k)
    shards = balanced_row_shards(seeds_h, N, N, world) if args.balance else shard_rows(N, world)
    r0, r1 = shards[rank]
    fake = None
    if args.fake_shard and world == 1:
        fr, fn = (int(x) for x in args.fake_shard.split("/"))
        shards = balanced_row_shards(seeds_h, N, N, fn) if args.balance else shard_rows(N, fn)
        r0, r1 = shards[fr]
        fake = (fr, fn)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maps(r0, r1):
        sig = device_signal_rows(N, r0, r1, torch, device)
        if mode == "cvt":
            return dict(density=(sig / 1.0) ** 2)
        return dict(signal=sig, noise=torch.ones_like(sig))

    seeds = torch.from_numpy(seeds_h).to(device)
    eng = Engine(local)
    drv = ShardedLloyd(eng) if world > 1 else None

    def setup(e, r0, r1, d=None):
        m = maps(r0, r1)
        e.set_image(y0=float(r0), **m)
        if (r1 - r0) != N:
            e.set_domain(0.0, 0.0, float(N), float(N))      # the generator table spans the full image
        (d.set_generators(seeds) if d is not None else e.set_generators(seeds))

    def step(n=1):
        if drv is not None:
            drv.step(n)
        elif fake is not None:
            # a lone shard must not move the generators (bins without pixels in this shard would
            # collapse to the origin): accumulate only
            for _ in range(n):
                eng.accumulate()
        else:
            eng.step(n)

    setup(eng, r0, r1, drv)
    rebalance = None
    if drv is not None and args.balance and not args.no_rebalance:
        # the cost model behind the cuts is good to a few per cent (and GPUs of one box differ by as
        # much): measure one round of accumulates, scale every shard's modelled cost to its measured
        # share, cut again, set up again -- all before the warm-up, nothing of it is timed
        for _ in range(3):                        # (the later cuts take out what the first left: one tile row at a time)
            t_acc = drv.measure_accumulate(n=6, skip=args.warmup)      # the iterations the timed region will see
            if not t_acc or max(t_acc) <= 1.01 * (sum(t_acc) / len(t_acc)):
                break
            new_shards = balanced_row_shards(seeds_h, N, N, world, measured=(shards, t_acc))
            if rebalance is None:
                rebalance = {"accumulate_ms_before": [round(x, 4) for x in t_acc], "shard_rows_before": [list(x) for x in shards]}
            if new_shards == shards:
                break
            shards = new_shards
            r0, r1 = shards[rank]
            setup(eng, r0, r1, drv)
        drv.set_generators(seeds)               # the timed iterations start from the seeds either way
    launches0 = eng.api.launch_count()
    if drv is not None:
        # communicator set-up is not a step: let NCCL finish its lazy channel/buffer initialisation on
        # the real message (the moment vector) before the W warm-up steps
        for _ in range(10):
            drv.exchange_moments()
        barrier()
    step(args.warmup)
    # NVML set-up (nvmlInit enumerates every GPU of the box: milliseconds) happens before the barrier,
    # otherwise the other ranks' timed regions would include rank 0's set-up through the first all-reduce
    sampler = ClockSampler(local)
    if rank == 0 and not args.no_clocks:
        sampler.start()
    barrier()
    sampler.active = True
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.api.launch_count()
    ev0.record()
    t_enq = time.perf_counter()
    step(args.steps)
    enqueue_ms = 1e3 * (time.perf_counter() - t_enq) / args.steps      # host time to enqueue one step
    ev1.record()
    barrier()
    gpu_launches = eng.api.launch_count() - l0
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if (rank == 0 and not args.no_clocks) else None
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = (float(N) * N) / (ms_step * 1e-3) / 1e9
    if fake is not None:
        value = (float(r1 - r0) * N) / (ms_step * 1e-3) / 1e9
    conv, _ = eng.status()

    if args.trace_steps:
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        barrier()
        evs[0].record()
        for i in range(args.steps):
            step(1)
            evs[i + 1].record()
        barrier()
        per = [round(evs[i].elapsed_time(evs[i + 1]), 3) for i in range(args.steps)]
        sys.stderr.write("rank %d per-step ms: %s\n" % (rank, per))

    # ---- dominant kernel, timed live with CUDA events on its own stream (second pass)
    eng.set_option(_capi.OPT_KERNEL_TIMING, 1)
    eng.kernel_timing()
    step(min(args.steps, 64))
    torch.cuda.synchronize()
    k_ms, k_n = eng.kernel_timing()
    eng.set_option(_capi.OPT_KERNEL_TIMING, 0)
    peak, peak_src = measured_peaks()
    roof = None
    if k_n > 0:
        k_ms_avg = k_ms / k_n
        alg_bytes = BYTES_PER_PX[mode] * float(r1 - r0) * N          # per launch (this rank's rows)
        per_rank = None
        if world > 1:
            # every rank's kernel time and bytes: the line reports the SLOWEST rank (the one the step
            # waits for) and the aggregate over the machine, not rank 0's shard
            t = torch.tensor([k_ms_avg, alg_bytes], dtype=torch.float64, device=device)
            allk = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allk, t)
            per_rank = [(float(x[0].item()), float(x[1].item())) for x in allk]
            slow = max(range(world), key=lambda i: per_rank[i][0])
            k_ms_avg, alg_bytes = per_rank[slow]
        ach = alg_bytes / (k_ms_avg * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "kernel": "k_image_main", "kernel_ms": k_ms_avg, "kernel_share_of_step": k_ms_avg / ms_step,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "step_frac_of_roofline": (BYTES_PER_PX[mode] * float(N) * N / world / (ms_step * 1e-3) / 1e9) / peak}
        if per_rank is not None:
            tot_b = sum(b for _, b in per_rank)
            t_max = max(ms for ms, _ in per_rank)
            roof["rank"] = "slowest of %d (rank %d)" % (world, slow)
            roof["aggregate_frac"] = (tot_b / world / (t_max * 1e-3) / 1e9) / peak
            roof["per_rank_kernel_ms"] = [round(ms, 4) for ms, _ in per_rank]
            roof["per_rank_frac"] = [round(b / (ms * 1e-3) / 1e9 / peak, 4) for ms, b in per_rank]
        tf = os.path.join(ROOT, "profiles", "traffic.json")      # dram bytes per launch from the ncu --set full capture
        if os.path.exists(tf):
            try:
                roof["traffic"] = json.load(open(tf)).get("%s_%s" % (wl_key, mode))
            except Exception:
                pass

    # ---- N > 1: where the iteration goes on every rank (accumulate | all-reduce | update), stderr only
    breakdown = None
    if world > 1:
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(8)]
        for e4 in ev:
            e4[0].record(); eng.accumulate(); e4[1].record()
            drv.exchange_moments(); e4[2].record()
            eng.update(); e4[3].record()
        torch.cuda.synchronize()
        acc = float(np.mean([e4[0].elapsed_time(e4[1]) for e4 in ev[2:]]))
        arr = float(np.mean([e4[1].elapsed_time(e4[2]) for e4 in ev[2:]]))
        upd = float(np.mean([e4[2].elapsed_time(e4[3]) for e4 in ev[2:]]))
        t = torch.tensor([acc, arr, upd, float(r1 - r0)], dtype=torch.float64, device=device)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        breakdown = [dict(rank=i, rows=int(x[3].item()), accumulate_ms=round(x[0].item(), 4),
                          allreduce_wait_ms=round(x[1].item(), 4), update_ms=round(x[2].item(), 4))
                     for i, x in enumerate(allt)]
        if rank == 0:
            sys.stderr.write("per-rank breakdown: %s\n" % json.dumps(breakdown))

    # ---- end to end through the public API with pinned host buffers
    e2e = None
    if not args.no_e2e:
        m = maps(r0, r1)
        host_in = {kk: v.cpu().pin_memory() for kk, v in m.items()}
        del m
        seeds_pin = torch.from_numpy(seeds_h).pin_memory()
        lab_pin = torch.empty((r1 - r0, N), dtype=torch.int32).pin_memory()
        node_pin = torch.empty((k, 2), dtype=torch.float64).pin_memory()
        e2 = Engine(local)
        d2 = ShardedLloyd(e2) if world > 1 else None

        def e2e_step():
            e2.set_image(y0=float(r0), **host_in)          # H2D of this step's maps (pinned -> device)
            if (r1 - r0) != N:
                e2.set_domain(0.0, 0.0, float(N), float(N))
            (d2.set_generators(seeds_pin) if d2 is not None else e2.set_generators(seeds_pin))   # H2D of the generator table
            (d2.step(1) if d2 is not None else e2.step(1))
            lab_pin.copy_(e2.labels(), non_blocking=True)   # D2H of the results
            node_pin.copy_(e2.generators(), non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(3):
            e2e_step()
        barrier()
        n_e2e = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        h2d = sum(v.numel() * 8 for v in host_in.values()) + seeds_pin.numel() * 8
        d2h = lab_pin.numel() * 4 + node_pin.numel() * 8
        e2e = {"value": (float(N) * N) / dt / 1e9, "unit": UNIT, "ms_per_step": dt * 1e3,
               "h2d_bytes_per_step": int(h2d * world), "d2h_bytes_per_step": int(d2h * world),
               "call": "Engine.set_image(pinned host maps) + set_generators + step(1) + labels()/generators() to pinned host"}
        # One GPU: the same steps double-buffered -- two Engine contexts on two CUDA streams, so that the
        # upload of step i+1 overlaps the download of step i (PCIe is full duplex).  Every step still
        # uploads its inputs and downloads its results; this is the throughput a user gets when
        # batches are pipelined.  (With N > 1 the two pipelines' collectives would interleave: serial.)
        if world == 1:
            eb = [e2, Engine(local)]
            sb = [torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)]
            lab_b = [lab_pin, torch.empty((r1 - r0, N), dtype=torch.int32).pin_memory()]
            node_b = [node_pin, torch.empty((k, 2), dtype=torch.float64).pin_memory()]

            def piped_step(i):
                e, st = eb[i & 1], sb[i & 1]
                with torch.cuda.stream(st):
                    e.set_image(y0=float(r0), **host_in)
                    e.set_generators(seeds_pin)
                    e.step(1)
                    lab_b[i & 1].copy_(e.labels(), non_blocking=True)
                    node_b[i & 1].copy_(e.generators(), non_blocking=True)

            for i in range(4):
                piped_step(i)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(n_e2e):
                piped_step(i)
            torch.cuda.synchronize()
            dtp = (time.perf_counter() - t0) / n_e2e
            e2e["serial_value"], e2e["serial_ms_per_step"] = e2e["value"], e2e["ms_per_step"]
            e2e["value"], e2e["ms_per_step"] = (float(N) * N) / dtp / 1e9, dtp * 1e3
            e2e["call"] += "; double-buffered over two Engine contexts / CUDA streams (serial_* = one context, one stream)"
            eb[1].close()
            del lab_b, node_b
        e2.close()
        del host_in, lab_pin

    # ---- N > 1: the same map on ONE GPU (rank 0), so the speed-up is apples to apples -- and an untimed
    #      parity check of the sharded path against it: after T identical iterations from the same
    #      seeds the generator tables agree to 1e-9 px, rank 0's whole shard and every rank's first
    #      and last row carry the same labels, and all ranks hold bit-identical tables
    single = None
    parity = None
    if world > 1 and not args.no_single:
        T_CHECK = 3
        barrier()
        drv.set_generators(seeds)
        drv.step(T_CHECK)
        g_sh = drv.generators()
        lab_sh = drv.local_labels()
        edge = torch.stack((lab_sh[0], lab_sh[-1])).contiguous()               # this rank's boundary rows
        edges = [torch.empty_like(edge) for _ in range(world)] if rank == 0 else None
        dist.gather(edge, edges, dst=0)
        h = (g_sh.view(torch.int64) * torch.arange(1, g_sh.numel() + 1, device=device).view_as(g_sh)).sum().view(1)
        hs = [torch.zeros_like(h) for _ in range(world)]
        dist.all_gather(hs, h)
        identical = all(int(x.item()) == int(hs[0].item()) for x in hs)
        if rank == 0:
            e1 = Engine(local)
            setup(e1, 0, N)
            e1.step(T_CHECK)
            g1 = e1.generators()
            lab1 = e1.labels()
            dg = float((g1 - g_sh).abs().max().item())
            shard_equal = bool(torch.equal(lab1[r0:r1], lab_sh))
            edges_equal = all(bool(torch.equal(edges[i][0], lab1[shards[i][0]])) and
                              bool(torch.equal(edges[i][1], lab1[shards[i][1] - 1])) for i in range(world))
            parity = {"ok": bool(dg <= 1e-9 and shard_equal and edges_equal and identical), "iterations": T_CHECK,
                      "max_abs_generator_diff_px": dg, "rank0_shard_labels_equal": shard_equal,
                      "boundary_rows_labels_equal": edges_equal, "tables_bit_identical_on_all_ranks": identical}
            del lab1, g1
            e1.set_generators(seeds)
            e1.step(args.warmup)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); e1.step(args.steps); b.record(); torch.cuda.synchronize()
            ms1 = a.elapsed_time(b) / args.steps
            single = {"ms_per_step": ms1, "value": (float(N) * N) / (ms1 * 1e-3) / 1e9, "unit": UNIT,
                      "speedup": ms1 / ms_step}
            e1.close()
        del lab_sh, g_sh
        barrier()

    # ---- N = 1: the other mode of the same workload and a seconds-long run
    modes = None
    sustained = None
    if world == 1 and fake is None and not args.no_extras:
        other = "sn" if mode == "cvt" else "cvt"
        eo = Engine(local)
        sig = device_signal_rows(N, 0, N, torch, device)
        if other == "sn":
            noise_o = torch.ones_like(sig)
            eo.set_image(signal=sig, noise=noise_o)
        else:
            dens_o = sig ** 2
            eo.set_image(density=dens_o)
        eo.set_generators(seeds)
        eo.step(args.warmup)
        eo.set_option(_capi.OPT_KERNEL_TIMING, 1)
        eo.kernel_timing()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eo.step(args.steps); b.record(); torch.cuda.synchronize()
        ko_ms, ko_n = eo.kernel_timing()
        eo.set_option(_capi.OPT_KERNEL_TIMING, 0)
        mso = a.elapsed_time(b) / args.steps
        bo = BYTES_PER_PX[other] * float(N) * N
        modes = {other: {"ms_per_step": mso, "value": float(N) * N / (mso * 1e-3) / 1e9, "unit": UNIT,
                         "bytes_per_px": BYTES_PER_PX[other], "kernel_ms": ko_ms / max(ko_n, 1),
                         "kernel_frac": bo / (ko_ms / max(ko_n, 1) * 1e-3) / 1e9 / peak,
                         "step_frac_of_roofline": bo / (mso * 1e-3) / 1e9 / peak}}
        eo.close()
        del eo, sig
        # sustained: >= 2 s of back-to-back iterations (chunks of 500, restarted from the seeds so that
        # the run never reaches the convergence latch), clocks sampled throughout
        samp2 = ClockSampler(local)
        if not args.no_clocks:
            samp2.start()
        eng.set_generators(seeds)
        eng.step(args.warmup)
        torch.cuda.synchronize()
        samp2.active = True
        tot_ms, tot_it = 0.0, 0
        while tot_ms < args.sustain_s * 1e3 and tot_it < 200000:
            eng.set_generators(seeds)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.step(500); b.record(); torch.cuda.synchronize()
            if eng.status()[0]:
                break
            tot_ms += a.elapsed_time(b); tot_it += 500
        ck = samp2.stop() if not args.no_clocks else None
        if tot_it:
            sustained = {"seconds": tot_ms * 1e-3, "iterations": tot_it, "ms_per_step": tot_ms / tot_it,
                         "value": float(N) * N / (tot_ms / tot_it * 1e-3) / 1e9, "unit": UNIT,
                         "step_frac_of_roofline": BYTES_PER_PX[mode] * float(N) * N / (tot_ms / tot_it * 1e-3) / 1e9 / peak,
                         "clocks": ck}

    # ---- CPU baseline: the compiled reference on a bounded slab (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rows = max(16, int(1.5e6 * 10.0 // N))            # ~20 s: two iterations over ~15 Mpix in total
        dt, npx, kind, rr = time_reference(N, k, min(rows, N), 2)
        cpu = {"value": npx / dt / 1e9, "unit": UNIT, "cores": 1, "kind": kind, "s_per_iter_sample": dt,
               "s_per_iter_full_map_extrapolated": dt * (float(N) * N / npx), "host_cpus": os.cpu_count(),
               "sample": "rows %d..%d of the %dx%d map (%d px) x all %d generators, lloyd(max_iters=0) = 2 iterations"
                         % (rr[0], rr[1], N, N, npx, k)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl["name"], "pixels": N * N, "generators": k, "mode": mode,
                           "bytes_per_px": BYTES_PER_PX[mode], "parallelism": ("rows x%d (%s) + allreduce(moments)" % (world, ("cost-balanced, re-cut from measured accumulate times" if rebalance else "cost-balanced") if args.balance else "equal")) if world > 1 else "1 GPU",
                           "exchange": drv.exchange if drv is not None else None,
                           "shard_rows": [list(x) for x in shards] if world > 1 or fake else None,
                           "rebalanced_from": rebalance,
                           "l2": "inputs (%.0f MB per GPU) larger than L2 (126 MB); no flush needed" % (
                               (BYTES_PER_PX[mode] - 4) * (r1 - r0) * N / 1e6),
                           "converged_during_run": bool(conv), "fake_shard": args.fake_shard,
                           "host_enqueue_ms_per_step": round(enqueue_ms, 4)},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(gpu_launches), "roofline": roof, "cpu_baseline": cpu}
        if single is not None:
            line["single_gpu_same_workload"] = single
        if parity is not None:
            line["parity_check"] = parity
        if modes is not None:
            line["modes"] = modes
        if sustained is not None:
            line["sustained"] = sustained
        if breakdown is not None:
            line["per_rank_breakdown"] = breakdown
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line))
        sys.stdout.flush()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
