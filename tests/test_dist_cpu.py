"""Host-side multi-GPU logic on CPU with the gloo backend, world_size 2: tile sharding + final
gather of the continent predictor and the data-parallel gradient all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_rank_tile_ranges_partition_the_plan():
    from deepbedmap_b200.tiler import max_tiles_per_rank, rank_row_band, rank_tile_range, tile_plan
    plan = tile_plan()
    for world in (1, 2, 3, 4, 8):
        covered = []
        for r in range(world):
            a, b = rank_tile_range(len(plan), r, world)
            assert 0 <= b - a <= max_tiles_per_rank(len(plan), world)
            covered += list(range(a, b))
            r0, r1 = rank_row_band(plan, r, world)
            assert all(r0 <= plan[i][0] and plan[i][1] <= r1 for i in range(a, b))
        assert covered == list(range(len(plan)))
    # 8 ranks: every band is a fraction of the 4500-row grid (what each rank uploads)
    bands = [rank_row_band(plan, r, 8) for r in range(8)]
    assert max(b - a for a, b in bands) < 4500 // 2


def _fake_tile(i, hh, ww):
    yy, xx = np.meshgrid(np.arange(hh), np.arange(ww), indexing="ij")
    return (i * 1000.0 + yy * 0.5 + xx * 0.25).astype(np.float32)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deepbedmap_b200 import tiler, train
        final, ary, pad = (120, 200), (40, 40), (3, 3)
        plan = tiler.tile_plan(final, ary, ary, pad)
        a, b = tiler.rank_tile_range(len(plan), rank, world)
        results = torch.full((tiler.max_tiles_per_rank(len(plan), world), ary[0], ary[1]), float("nan"))
        for i in range(a, b):
            _, _, _, _, ys, ye, xs, xe = plan[i]
            results[i - a, :ye - ys, :xe - xs] = torch.from_numpy(_fake_tile(i, ye - ys, xe - xs))

        def place(src, canvas, ys, xs, hh, ww):
            canvas[ys:ys + hh, xs:xs + ww] = src[:hh, :ww]

        canvas = tiler.gather_and_assemble(results, plan, final, ary, rank, world,
                                           lambda: torch.full(final, float("nan")), place)
        if rank == 0:
            ref = np.full(final, np.nan, np.float32)
            for i, (_, _, _, _, ys, ye, xs, xe) in enumerate(plan):
                ref[ys:ye, xs:xe] = _fake_tile(i, ye - ys, xe - xs)
            q.put(("tiler", bool(np.array_equal(canvas.numpy(), ref, equal_nan=True))))
        else:
            assert canvas is None

        class Link:  # only what allreduce_grads touches
            flat_grad = torch.arange(10, dtype=torch.float32) * (rank + 1)

        scale = train.allreduce_grads(Link)
        expect = torch.arange(10, dtype=torch.float32) * 3  # (1 + 2)
        ok = bool(torch.equal(Link.flat_grad, expect)) and scale == 0.5
        # bucketed variant: buckets arrive in backward order (tail of the buffer first)
        class Link2:
            flat_grad = torch.arange(12, dtype=torch.float32) * (rank + 1)

        red = train.GradBucketReducer(Link2)
        for lo, hi in ((9, 12), (4, 9), (0, 4)):
            red.bucket(lo, hi)
        ok = ok and red.finish() == 0.5 and bool(torch.equal(Link2.flat_grad, torch.arange(12, dtype=torch.float32) * 3))
        bad = train.GradBucketReducer(Link2)
        bad.bucket(0, 4)
        try:
            bad.finish()
            ok = False
        except RuntimeError:
            pass
        if rank == 0:
            q.put(("allreduce", ok))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_and_allreduce_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = dict(q.get(timeout=10) for _ in range(2))
    assert got == {"tiler": True, "allreduce": True}


def test_single_process_defaults():
    from deepbedmap_b200 import tiler, train

    class Link:
        flat_grad = torch.ones(4)

    assert train.allreduce_grads(Link) == 1.0 and torch.equal(Link.flat_grad, torch.ones(4))
    assert tiler.rank_tile_range(396, 0, 1) == (0, 396)
