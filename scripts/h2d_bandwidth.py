"""Aggregate pinned host -> device bandwidth of N ranks uploading at the same time (one process per GPU, torchrun):
the bound of the end-to-end continent run, which uploads 10.9 GB of input grids per continent. Prints per-rank and
aggregate GB/s for contiguous copies and for the tiler's strided 2-D column blocks."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl")
gb = 1.0
rows, cols = 2880, int(gb * 2**30 / 4 / 2880)
host = torch.empty(rows, cols, dtype=torch.float32, pin_memory=True).normal_()
dev = torch.empty_like(host, device="cuda")


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=5):
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    return reps * host.numel() * 4 / dt / 1e9


def blocks():   # torch's strided copy_ (what a caller without the library would write)
    nb = 11
    for k in range(nb):
        a, b = cols * k // nb, cols * (k + 1) // nb
        dev[:, a:b].copy_(host[:, a:b], non_blocking=True)


def blocks2d():   # cudaMemcpy2DAsync per block: what deepbedmap_b200.tiler.StreamedGrids does for pinned grids
    import sys
    sys.path.insert(0, ".")
    from deepbedmap_b200 import ops
    nb = 11
    st = torch.cuda.current_stream().cuda_stream
    for k in range(nb):
        a, b = cols * k // nb, cols * (k + 1) // nb
        ops.call("dbm_copy2d_async", dev.data_ptr() + 4 * a, cols * 4, host.data_ptr() + 4 * a, cols * 4, (b - a) * 4, rows, st)


res = torch.tensor([timed(lambda: dev.copy_(host, non_blocking=True)), timed(blocks), timed(blocks2d)], device="cuda")
if world > 1:
    allr = [torch.empty_like(res) for _ in range(world)]
    dist.all_gather(allr, res)
else:
    allr = [res]
if rank == 0:
    c = [float(r[0]) for r in allr]
    b = [float(r[1]) for r in allr]
    d = [float(r[2]) for r in allr]
    print(f"{world} rank(s) uploading 1 GiB each concurrently: contiguous {sum(c):.1f} GB/s aggregate "
          f"({min(c):.1f} .. {max(c):.1f} per rank); 11 column blocks: torch strided copy_ {sum(b):.1f} GB/s aggregate "
          f"({min(b):.1f} .. {max(b):.1f} per rank), cudaMemcpy2DAsync {sum(d):.1f} GB/s aggregate "
          f"({min(d):.1f} .. {max(d):.1f} per rank)")
if world > 1:
    dist.destroy_process_group()
