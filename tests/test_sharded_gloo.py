"""
The N > 1 host path on the CPU: two processes, `gloo` backend, each driving the emulator build of
the library through tess_b200.sharded.ShardedLloyd (row shards, one all-reduce of the moment vector
per iteration).  Results must equal the reference trajectory of the unsharded problem.
"""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, golden
from tess_b200.sharded import shard_rows


def test_shard_rows_cover_and_align():
    for H, world in ((8192, 8), (32768, 8), (100, 3), (64, 2), (5, 4), (200, 2)):
        r = shard_rows(H, world)
        assert r[0][0] == 0 and r[-1][1] == H and len(r) == world
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        if H >= 64 * world:
            assert all(a % 64 == 0 for a, _ in r) and all(b > a for a, b in r)


def test_balanced_row_shards():
    from tess_b200.sharded import balanced_row_shards
    rng = np.random.default_rng(0)
    H = W = 4096
    g = np.column_stack((rng.normal(W / 2, 300, 20000), rng.normal(H / 2, 300, 20000)))
    r = balanced_row_shards(g, H, W, 8)
    assert r[0][0] == 0 and r[-1][1] == H and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    assert all(a % 64 == 0 for a, _ in r) and all(b > a for a, b in r)
    heights = [b - a for a, b in r]
    assert min(heights[3:5]) < heights[0]            # central shards are thinner than the outer ones
    assert balanced_row_shards(g, 100, W, 4) == shard_rows(100, 4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    from harness import EmuEngine
    from tess_b200.sharded import ShardedLloyd, shard_rows as sr
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = np.load(os.path.join(ROOT, "tests", "golden", "lloyd_img64.npz"))
        dens = g["density"]
        r0, r1 = sr(dens.shape[0], world)[rank]
        sh = ShardedLloyd(EmuEngine())
        sh.set_image_rows(density=dens[r0:r1], row0=r0, total_rows=dens.shape[0])
        sh.set_generators(g["seeds"])
        # teacher-free trajectory, checked every iteration on every rank
        T = int(g["n_iters"])
        ok = True
        for t in range(3):
            sh.step(1)
            ok &= np.array_equal(sh.local_labels(), g["labels"][t].reshape(64, 64)[r0:r1].astype(np.int32))
            ok &= np.abs(sh.generators() - g["nodes"][t]).max() < 1e-9
        sh.set_generators(g["seeds"])
        conv, nit = sh.run(300)
        full = sh.gather_labels(0)
        # no peer memory on the CPU: the exchange must have fallen back to the backend's all_reduce
        ok &= sh.exchange.startswith("backend all_reduce (symmetric memory unavailable")
        res = dict(rank=rank, ok=bool(ok), conv=conv, nit=nit, nodes=sh.generators(),
                   mom=sh.moments(), labels=None if full is None else full.numpy())
        q.put(res)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_row_sharded_lloyd_matches_reference():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    g = golden("lloyd_img64.npz")
    res.sort(key=lambda r: r["rank"])
    for r in res:
        assert r["ok"] and r["conv"] and r["nit"] == int(g["n_iters"])
        assert np.abs(r["nodes"] - g["nodes"][-1]).max() < 1e-9
    # all ranks hold bit-identical sums and nodes after the all-reduce
    assert np.array_equal(res[0]["nodes"], res[1]["nodes"]) and np.array_equal(res[0]["mom"], res[1]["mom"])
    assert np.array_equal(res[0]["labels"].ravel(), g["labels"][-1].astype(np.int32))
    assert res[0]["mom"][:, 0].sum() == 64 * 64


def test_balanced_row_shards_and_recut_from_measured_times():
    """Host logic of the row sharding: cuts on tile rows, cover the image, follow the generator density;
    a re-cut from measured accumulate times shrinks the shards that ran long and leaves a balanced
    measurement alone."""
    import numpy as np
    from tess_b200.sharded import balanced_row_shards, shard_rows
    rng = np.random.default_rng(3)
    H = W = 4096
    g = np.clip(rng.normal(H / 2, H / 8, (20000, 2)), 0, H - 1)
    cuts = balanced_row_shards(g, H, W, 8)
    assert cuts[0][0] == 0 and cuts[-1][1] == H and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
    assert all(r0 % 64 == 0 and r1 > r0 for r0, r1 in cuts)
    rows = [r1 - r0 for r0, r1 in cuts]
    assert rows[0] > rows[3] and rows[7] > rows[4]                      # sparse outskirts get more rows
    assert balanced_row_shards(g, H, W, 1) == shard_rows(H, 1)
    same = balanced_row_shards(g, H, W, 8, measured=(cuts, [1.0] * 8))
    # equal measured times: the modelled cost is rescaled to equal shares -> the cuts stay within a tile row
    assert all(abs(a[1] - b[1]) <= 64 for a, b in zip(cuts, same))
    slow0 = balanced_row_shards(g, H, W, 8, measured=(cuts, [1.3] + [1.0] * 7))
    assert slow0[0][1] < cuts[0][1]                                      # rank 0 ran long: fewer rows for it
    assert slow0[-1][1] == H and all(a[1] == b[0] for a, b in zip(slow0, slow0[1:]))
    # a density-independent model gives equal rows
    assert balanced_row_shards(g, H, W, 8, model=(1.0, 0.0, 1.0)) == shard_rows(H, 8)
