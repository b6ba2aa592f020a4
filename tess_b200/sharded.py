"""
tess_b200.sharded -- the Lloyd / WVT iteration over pixel rows sharded across GPUs.

One process per GPU (torch.distributed, NCCL over NVLink).  Rank r owns a contiguous block of
image rows (its slice of the density / signal / noise maps and of the label map); the generator
table is replicated and every rank rebuilds the same index.  Per iteration exactly one collective
runs: an all-reduce (sum, fp64) of the packed per-bin moment vector [k*M | changed-flag], which is
the only data that couples pixels across GPUs.  All ranks then hold bit-identical
sums, so they take the same node update and the same convergence decision without a broadcast.

The reference has no distributed path; the single-process semantics being reproduced are those of
tess/lloyd.pyx:47-90.
"""
from ._capi import TessNonFinite


def shard_rows(H, world, align=64):
    """Row ranges [(r0, r1), ...] of `world` contiguous shards of an H-row image.  Boundaries fall
    on multiples of `align` (the kernels' 64-row super-tiles) where H allows it; a rank may get an
    empty range only if H < world."""
    blocks = (H + align - 1) // align
    if blocks >= world:
        cuts = [min(H, ((blocks * r) // world) * align) for r in range(world + 1)]
    else:
        cuts = [(H * r) // world for r in range(world + 1)]
    cuts[-1] = H
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


# time per pixel of one accumulate as a function of the local generator density rho (generators per
# pixel): a + b rho^p picoseconds (see balanced_row_shards)
COST_MODEL = (2.56, 39.1, 0.40)


def balanced_row_shards(generators, H, W, world, x0=0.0, y0=0.0, align=64, model=None, measured=None):
    """Row ranges whose estimated cost is equal: the time per pixel of one accumulate (list building +
    image kernel) grows with the local generator density rho (generators per pixel) like
    2.56 + 39.1 rho^0.40 ps/px -- a least-squares fit to the per-rank accumulate times of the C4 map
    on 8 GPUs under two different shardings -- so equal-height shards of a centrally concentrated map
    would leave the central ranks 2x slower than the outer ones.  Checked against the final round-2
    kernels on one GPU (tools/shard_cost_probe.py: the eight shards of C4 take 0.46-0.49 ms each,
    max / mean = 1.027; a refit to 96 strips of 512 and 2048 rows, 2.28 + 63.5 rho^0.50, predicts the
    same cuts to within one tile row, i.e. the model is at the noise limit of the measurements).
    `generators` is the (k, 2) table (host array); boundaries fall on multiples of `align` rows;
    `model` = (a, b, p) overrides COST_MODEL.
    `measured` = (shards, times): row ranges in use and the accumulate time every rank measured with
    them (ShardedLloyd.measure_accumulate).  The model's cost of each of those shards is then scaled to
    its measured share before the rows are cut again -- the model is good to a few per cent, a GPU of
    the box may be a few per cent slower than its peers, and the slowest rank sets the iteration."""
    ca, cb, cp = model if model is not None else COST_MODEL
    import numpy as np
    g = np.asarray(generators, dtype=np.float64)
    nty, ntx = (H + align - 1) // align, (W + align - 1) // align
    if world <= 1 or nty < world:
        return shard_rows(H, world, align)
    ix = np.clip(((g[:, 0] - x0) // align).astype(np.int64), 0, ntx - 1)
    iy = np.clip(((g[:, 1] - y0) // align).astype(np.int64), 0, nty - 1)
    cnt = np.bincount(iy * ntx + ix, minlength=nty * ntx).reshape(nty, ntx).astype(np.float64)
    tile_px = float(align * align)
    cost = (tile_px * (ca + cb * (cnt / tile_px) ** cp)).sum(axis=1)     # per tile row
    if measured is not None:
        old_shards, times = measured
        times = np.asarray(times, dtype=np.float64)
        if len(old_shards) == len(times) and np.all(times > 0):
            for (a0, a1), t in zip(old_shards, times):
                t0, t1 = a0 // align, (a1 + align - 1) // align
                c = cost[t0:t1].sum()
                if c > 0:
                    cost[t0:t1] *= (t / times.sum()) / (c / cost.sum() if cost.sum() > 0 else 1.0)
    cum = np.concatenate(([0.0], np.cumsum(cost)))
    cuts = [0]
    for r in range(1, world):
        t = int(np.searchsorted(cum, cum[-1] * r / world))
        t = min(max(t, cuts[-1] // align + 1), nty - (world - r))              # at least one tile row each
        cuts.append(t * align)
    cuts.append(H)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


class ShardedLloyd(object):
    """Drives one `Engine` per rank; `dist` is torch.distributed (already initialised)."""

    def __init__(self, engine, group=None, exchange="auto", renumber=True):
        """`renumber`: number the generators in spatial (row-major 64-px cell) order inside the engine,
        so that the bins a rank's rows touch are one short range of ids and the exchange can skip the
        rest (tess_peer_allreduce_sparse); `generators()`, `moments()`, `local_labels()` and
        `gather_labels()` translate back to the caller's numbering.  An exact distance tie then goes to
        the lower SPATIAL index, not the lower caller index (ties need structured input: generators on
        a lattice).
        `exchange`: how the moment vector is summed over the ranks each iteration --
        "peer": this library's own kernel over symmetric (peer-mapped) memory, the reduction done in
        the NVSwitch when a multicast mapping exists (tess_peer_allreduce); "nccl": one
        `all_reduce`; "auto": "peer" on CUDA when symmetric memory can be set up, else "nccl".
        `self.exchange` says what is in use (and why, after a fallback)."""
        import os
        import torch
        import torch.distributed as dist
        self.eng = engine
        self.dist = dist
        self.torch = torch
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._view = None
        self._want = os.environ.get("TESS_EXCHANGE", exchange)
        self._renumber = bool(renumber) and os.environ.get("TESS_NO_RENUMBER", "") != "1"
        self._perm = None        # internal id -> caller id (None: identity)
        self._inv = None         # caller id -> internal id
        self._sparse = False
        self._owner = False
        self._peer = None
        self.exchange = "backend all_reduce (NCCL on GPUs)" if self._want == "nccl" else "undecided"

    # ------------------------------------------------------------------ inputs
    def set_image_rows(self, density=None, signal=None, noise=None, row0=0, x0=0.0, y0=0.0, total_rows=None):
        """This rank's rows of the maps; `row0` is the global index of its first row, `total_rows` the
        height of the full image (the replicated generator table is binned over the full image)."""
        self.eng.set_image(density=density, signal=signal, noise=noise, x0=x0, y0=float(y0) + float(row0))
        if total_rows is not None:
            W = self.eng.shape[1]
            self.eng.set_domain(float(x0), float(y0), float(x0) + W, float(y0) + float(total_rows))
        self._view = None

    def _spatial_order(self, node):
        """Stable order of the generators by (row, column) of their 64-px cell: the same on every rank."""
        if isinstance(node, self.torch.Tensor):
            t = self.torch
            cy = t.floor(node[:, 1] / 64.0).clamp_(-2.0 ** 20, 2.0 ** 20).to(t.int64)
            cx = t.floor(node[:, 0] / 64.0).clamp_(-2.0 ** 20, 2.0 ** 20).to(t.int64)
            perm = t.argsort(cy * (1 << 22) + cx, stable=True)
            inv = t.empty_like(perm)
            inv[perm] = t.arange(perm.numel(), device=perm.device)
            return perm, inv
        import numpy as np
        node = np.asarray(node)
        cy = np.clip(np.floor(node[:, 1] / 64.0), -2.0 ** 20, 2.0 ** 20).astype(np.int64)
        cx = np.clip(np.floor(node[:, 0] / 64.0), -2.0 ** 20, 2.0 ** 20).astype(np.int64)
        perm = np.argsort(cy * (1 << 22) + cx, kind="stable")
        inv = np.empty_like(perm)
        inv[perm] = np.arange(perm.size)
        return perm, inv

    def set_generators(self, node_xy, scale=None):
        """Same table on every rank (replicated)."""
        self._perm = self._inv = None
        if self._renumber and self.world > 1:
            node = self.eng._to_dev(node_xy)
            self._perm, self._inv = self._spatial_order(node)
            node_xy = node[self._perm]
            if scale is not None:
                scale = self.eng._to_dev(scale)[self._perm]
        self.eng.set_generators(node_xy, scale)
        self._view = None
        if self._want != "nccl":
            self._setup_peer_exchange()

    def _setup_peer_exchange(self):
        """Moment vector in a symmetric allocation (torch.distributed._symmetric_memory: peer-mapped
        on every rank of the node, multicast-mapped when the fabric supports NVLS)."""
        need = 6 * self.eng.k + 2 + 34     # vector (6 moments at most) + tail + barrier flags and range slots of the sparse exchange
        if self._peer is not None and self._peer["n"] >= need:
            return
        if self._peer is not None:          # outgrown: back to a library-owned vector until replaced
            self.eng.set_moment_buffer(None, 0)
            self._peer = None
            self._view = None
        import ctypes
        import os
        gen = self.eng.generators()
        if self.world < 2 or self.world > 16 or not isinstance(gen, self.torch.Tensor):
            # the same on every rank (CPU tests with NumPy memory, a lone rank): no agreement needed
            why = "engine memory is not CUDA" if self.world >= 2 else "world size %d" % self.world
            if self._want == "peer":
                raise RuntimeError("peer exchange unavailable: " + why)
            self.exchange = "backend all_reduce (symmetric memory unavailable: %s)" % why
            return
        dev, err, peer = gen.device, None, None
        symm = buf = None
        try:
            import torch.distributed._symmetric_memory as symm
            buf = symm.empty(need, dtype=self.torch.float64, device=dev)
            buf.zero_()
        except Exception as e:      # no symmetric memory on this rank (driver / fabric support missing)
            err = str(e).splitlines()[0][:120] if str(e) else type(e).__name__
            buf = None
        # first vote: the rendezvous below is itself collective, so it is entered only when every
        # rank got its allocation (a rank that failed above would otherwise leave its peers waiting)
        ok = self.torch.tensor([1 if buf is not None else 0], dtype=self.torch.int32, device=dev)
        self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 1:
            try:
                grp = self.group if self.group is not None else self.dist.group.WORLD
                hdl = symm.rendezvous(buf, grp)
                ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in hdl.buffer_ptrs])
                mc = int(hdl.multicast_ptr) if getattr(hdl, "has_multicast_support", False) and hdl.multicast_ptr else 0
                mc_raw = mc
                # fp64 multimem operations are 8 bytes each (no vector form): measured 126 us per 32 MB
                # exchange at 8 GPUs against 141 us for the two-shot kernel and 147 us for NCCL, but
                # 128 us against 83 / 80 us at 2 GPUs -- the switch reduction pays from 8 ranks on
                force = os.environ.get("TESS_PEER_NO_MC", "")
                if force == "1" or (force != "0" and self.world < 8):
                    mc = 0
                peer = dict(buf=buf, hdl=hdl, ptrs=ptrs, mc=mc, mc_raw=mc_raw, n=need)
            except Exception as e:
                err = str(e).splitlines()[0][:120] if str(e) else type(e).__name__
        else:
            err = err or "a peer rank could not allocate symmetric memory"
        # second vote, all ranks or none: the exchange is collective, so every rank takes part in it
        # whatever happened above
        ok = self.torch.tensor([1 if peer is not None else 0], dtype=self.torch.int32, device=dev)
        self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 1:
            self.eng.set_moment_buffer(peer["buf"], need)
            self._peer = peer
            self._view = None
            import os
            self._sparse = self._perm is not None and os.environ.get("TESS_PEER_DENSE", "") != "1"
            use_mc = self._sparse and peer["mc_raw"] and os.environ.get("TESS_SPARSE_MC", "0") == "1"
            self.eng.peer_set_multicast(peer["mc_raw"] if use_mc else 0)
            # owner-computed update: the owner of a slice broadcasts the new nodes of its bins (16 bytes
            # each) instead of the rows of totals (the library applies it to 4-moment rows without WVT
            # scale updates and falls back to the rows otherwise)
            self._owner = self._sparse and os.environ.get("TESS_OWNER_UPDATE", "1") != "0"
            from . import _capi
            self.eng.set_option(_capi.OPT_OWNER_UPDATE, 1 if self._owner else 0)
            if self._sparse:
                self.exchange = "own kernel, two-shot over NVLink peer memory, sparse reduce (spatially numbered bins)"
                if self._owner:
                    self.exchange += ", owner-computed nodes broadcast (16 B per bin; 4-moment rows)"
                if use_mc:
                    self.exchange += ", broadcast through the NVLS multicast mapping"
            else:
                self.exchange = ("own kernel, NVLS multicast reduction (multimem.ld_reduce/st)" if peer["mc"]
                                 else "own kernel, two-shot over NVLink peer memory")
            return
        err = err or "a peer rank could not set up symmetric memory"
        if self._want == "peer":
            raise RuntimeError("peer exchange unavailable: " + err)
        self.exchange = "backend all_reduce (symmetric memory unavailable: %s)" % err

    # ------------------------------------------------------------------ iteration
    def _moments(self):
        if self._view is None:
            v, _ = self.eng.moments_view()
            if not isinstance(v, self.torch.Tensor):          # NumPy-memory engine (CPU tests, gloo)
                v = self.torch.from_numpy(v)
            self._view = v
        return self._view

    def step(self, n=1):
        """n iterations: local accumulate -> all-reduce of the moment vector -> update."""
        for _ in range(int(n)):
            self.eng.accumulate()
            self.exchange_moments()
            self.eng.update()

    def measure_accumulate(self, n=3, skip=1):
        """Device time of one local accumulate on every rank (mean of n, after `skip` untimed ones; the
        exchange and update run in between as in a real iteration): the input of
        balanced_row_shards(measured=...).  Collective; returns the same list on every rank."""
        t = self.torch
        if not isinstance(self.eng.generators(), t.Tensor):
            return None
        ev = [(t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)) for _ in range(n + skip)]
        for a, b in ev:
            a.record()
            self.eng.accumulate()
            b.record()
            self.exchange_moments()
            self.eng.update()
        t.cuda.synchronize()
        mine = t.tensor([sum(a.elapsed_time(b) for a, b in ev[skip:]) / n], dtype=t.float64, device=self.eng.generators().device)
        allt = [t.zeros_like(mine) for _ in range(self.world)]
        self.dist.all_gather(allt, mine, group=self.group)
        return [float(x.item()) for x in allt]

    def exchange_moments(self):
        """The one cross-rank step: every rank ends with the bit-identical sum of the moment vectors."""
        if self._peer is None:
            self.dist.all_reduce(self._moments(), op=self.dist.ReduceOp.SUM, group=self.group)
            return
        p = self._peer
        n = self.eng.api.moments_ptr(self.eng.ctx)[1]
        if self._sparse:
            # one kernel: its own cross-rank barriers around the sparse reduce + broadcast (the range of
            # non-zero rows was published by the accumulate)
            self.eng.peer_allreduce_sparse(self.world, self.rank, p["ptrs"], n)
            return
        p["hdl"].barrier(channel=0)            # every rank's local sums are final (stream-ordered)
        self.eng.peer_allreduce(self.world, self.rank, p["ptrs"], p["mc"], n)
        p["hdl"].barrier(channel=1)            # every rank's slice has been broadcast

    def run(self, max_iters, check_every=4):
        """Reference stop rule (tess/lloyd.pyx:85-90) on the sharded problem.
        Returns (converged, n_iters); identical on every rank."""
        budget = max(int(max_iters) + 2, 1)
        start = self.eng.stats()[2]
        done, conv = 0, False
        while not conv and done < budget:
            chunk = min(int(check_every), budget - done)
            self.step(chunk)
            done = self.eng.stats()[2] - start
            conv, bad = self.eng.status()
            if bad:
                if done < budget:
                    raise TessNonFinite(-4, "a bin with pixels has zero total weight: generator became "
                                            "NaN/Inf after iteration %d" % done)
                return False, done
        return conv, done

    # ------------------------------------------------------------------ results
    def generators(self):
        g = self.eng.generators()
        return g if self._inv is None else g[self._inv]

    def local_labels(self):
        """Labels of this rank's rows, in the caller's numbering (they never leave the GPU unless
        asked for)."""
        lab = self.eng.labels()
        if self._perm is None:
            return lab
        if isinstance(lab, self.torch.Tensor):
            return self._perm[lab.long()].to(lab.dtype)
        return self._perm[lab].astype(lab.dtype)

    def moments(self):
        """Global per-bin moments [k, 6] (identical on all ranks after a step)."""
        m = self.eng.moments()
        if self._owner and self._peer is not None and self.eng.api.moments_ptr(self.eng.ctx)[2] == 4 \
                and not getattr(self.eng, "_wvt_update", False):
            # owner-computed update: the totals of a bin live in the copy of the rank that owns its slice
            # (rows [r * ceil(k / world), ...)); collect the slices
            k = self.eng.k
            per = (k + self.world - 1) // self.world
            lo, hi = min(per * self.rank, k), min(per * (self.rank + 1), k)
            mine = self.torch.zeros((per, m.shape[1]), dtype=m.dtype, device=m.device)
            mine[:hi - lo] = m[lo:hi]
            parts = [self.torch.empty_like(mine) for _ in range(self.world)]
            self.dist.all_gather(parts, mine, group=self.group)
            m = self.torch.cat(parts, dim=0)[:k]
        return m if self._inv is None else m[self._inv]

    def gather_labels(self, dst=0):
        """Full label map on rank `dst` (None elsewhere); rows are concatenated in rank order."""
        lab = self.local_labels()
        if not isinstance(lab, self.torch.Tensor):
            lab = self.torch.from_numpy(lab)
        sizes = [None] * self.world
        self.dist.all_gather_object(sizes, tuple(lab.shape), group=self.group)
        if self.rank == dst:
            parts = [self.torch.empty(s, dtype=lab.dtype, device=lab.device) for s in sizes]
        else:
            parts = None
        self.dist.gather(lab.contiguous(), parts, dst=dst, group=self.group)
        return self.torch.cat(parts, dim=0) if self.rank == dst else None
