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


def test_output_segments_partition_the_canvas():
    """Streaming host output: over all ranks, the rectangles each rank copies into the shared host DEM cover every
    canvas pixel exactly once (tile windows + the NaN frame), for the reference geometry and a small ragged one."""
    from deepbedmap_b200.tiler import output_segments, rank_tile_range, tile_plan, windows_abut
    for final, ary, pad in (((18000, 22000), (1000, 1000), (18, 18)), ((120, 200), (40, 40), (3, 3))):
        plan = tile_plan(final, ary, ary, pad)
        tiles_x = final[1] // ary[1]
        assert windows_abut(plan, tiles_x)
        for world in (1, 2, 3, 5, 8):
            cover = np.zeros((final[0] // 4, final[1] // 4), np.int32)   # all windows are multiples of 4 px
            for r in range(world):
                a, b = rank_tile_range(len(plan), r, world)
                for ys, ye, xs, xe, i, j in output_segments(plan, a, b, final, tiles_x):
                    assert a <= i < j <= b and i // tiles_x == (j - 1) // tiles_x
                    assert ys % 4 == 0 and ye % 4 == 0 and xs % 4 == 0 and xe % 4 == 0
                    # the segment contains the windows of its own tiles
                    for t in plan[i:j]:
                        assert ys <= t[4] and t[5] <= ye and xs <= t[6] and t[7] <= xe
                    cover[ys // 4:ye // 4, xs // 4:xe // 4] += 1
            assert cover.min() == 1 and cover.max() == 1, (final, world)
    # overlapping windows (stride < ary_shape) are refused by the streaming path
    assert not windows_abut(tile_plan((120, 200), (40, 40), (20, 20), (3, 3)), len(range(0, 200, 20)))


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

        # shared host DEM: every rank writes its own segments; rank 0 reads the assembled grid
        dem = tiler.HostDEM(final)
        segs = tiler.output_segments(plan, a, b, final, final[1] // ary[1])
        local = np.full(final, np.nan, np.float32)
        for i in range(a, b):
            _, _, _, _, ys, ye, xs, xe = plan[i]
            local[ys:ye, xs:xe] = _fake_tile(i, ye - ys, xe - xs)
        for ys, ye, xs, xe, _, _ in segs:
            dem.array[0, ys:ye, xs:xe] = local[ys:ye, xs:xe]
        dist.barrier()
        if rank == 0:
            q.put(("hostdem", bool(np.array_equal(dem.array[0], ref, equal_nan=True))))
        dem.close()

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
    got = dict(q.get(timeout=10) for _ in range(3))
    assert got == {"tiler": True, "hostdem": True, "allreduce": True}


def test_single_process_defaults():
    from deepbedmap_b200 import tiler, train

    class Link:
        flat_grad = torch.ones(4)

    assert train.allreduce_grads(Link) == 1.0 and torch.equal(Link.flat_grad, torch.ones(4))
    assert tiler.rank_tile_range(396, 0, 1) == (0, 396)
