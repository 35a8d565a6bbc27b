"""Tiled whole-continent predictor (deepbedmap.py:681-740), tiles sharded over GPUs.

Tile geometry is reproduced exactly (it is semantics: the crop borders decide where the trunk's
zero padding falls): 18 x 22 output tiles of 1000 x 1000 px, each predicted from a lowres crop
with an 18+1 px halo, the outer 72 px of every prediction cropped, the canvas pre-filled with NaN.

B200 design: the continent grids (10.9 GB) are uploaded ONCE and stay resident in HBM (the
reference re-uploads 14.4 GB of overlapping crops); crop + clip>=0 (deepbedmap.py:663-665,
715-722) is one small kernel per input, same-shape tiles are batched through the generator, and
predictions are placed on a device-resident canvas that is read back once.  With
torch.distributed initialised, each rank takes a contiguous run of tiles (so it only needs - and
only uploads - the band of grid rows under its tiles) and rank 0 gathers the per-rank result
stacks: the only collective.
"""
from __future__ import annotations

import os

from collections import OrderedDict
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch


def tile_plan(final_shape=(18000, 22000), ary_shape=(1000, 1000), stride=(1000, 1000), xtrapad=(18, 18)):
    """(y0, y1, x0, x1, ys, ye, xs, xe) per tile, in the reference's order (deepbedmap.py:700-732)."""
    plan = []
    for sy in range(0, final_shape[0], stride[0]):
        for sx in range(0, final_shape[1], stride[1]):
            y0 = max(0, (sy // 4) - xtrapad[0] - 1)
            y1 = min(final_shape[0] // 4, ((sy + ary_shape[0]) // 4) + xtrapad[0] + 1)
            x0 = max(0, (sx // 4) - xtrapad[1] - 1)
            x1 = min(final_shape[1] // 4, ((sx + ary_shape[1]) // 4) + xtrapad[1] + 1)
            plan.append((y0, y1, x0, x1, (y0 + xtrapad[0] + 1) * 4, (y1 - xtrapad[0] - 1) * 4,
                         (x0 + xtrapad[1] + 1) * 4, (x1 - xtrapad[1] - 1) * 4))
    return plan


# ---- host-side sharding logic (device independent; covered by gloo tests on CPU) -----------------
def rank_tile_range(n_tiles: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced run of tile indices owned by ``rank``."""
    base, rem = divmod(n_tiles, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def max_tiles_per_rank(n_tiles: int, world: int) -> int:
    return (n_tiles + world - 1) // world


def rank_row_band(plan, rank: int, world: int) -> Tuple[int, int]:
    """Lowres row range [r0, r1) of the input grids that the rank's tiles read."""
    a, b = rank_tile_range(len(plan), rank, world)
    if a == b:
        return 0, 0
    return min(t[0] for t in plan[a:b]), max(t[1] for t in plan[a:b])


def output_segments(plan, a: int, b: int, final_shape, tiles_x: int):
    """Host-DEM rectangles owned by the contiguous tile run [a, b): one (ys, ye, xs, xe) per tile row the run
    touches, covering the canvas windows of the run's tiles in that row and extended to the canvas border where the
    run holds the row's first / last tile (columns) or the first / last tile row (rows), so that the NaN frame the
    reference leaves around its product (deepbedmap.py:696, 731-736) travels with the data. Over all ranks the
    rectangles partition the canvas exactly (asserted by tests/test_dist_cpu.py) provided the tile windows abut,
    i.e. stride == ary_shape as in the reference (checked by ``windows_abut``)."""
    n_rows = len(plan) // tiles_x
    segs = []
    i = a
    while i < b:
        ty = i // tiles_x
        j = min(b, (ty + 1) * tiles_x)
        ys, ye = plan[i][4], plan[i][5]
        xs, xe = plan[i][6], plan[j - 1][7]
        if i % tiles_x == 0:
            xs = 0
        if j % tiles_x == 0:
            xe = final_shape[1]
        if ty == 0:
            ys = 0
        if ty == n_rows - 1:
            ye = final_shape[0]
        segs.append((ys, ye, xs, xe, i, j))
        i = j
    return segs


def windows_abut(plan, tiles_x: int) -> bool:
    """True when consecutive tiles' canvas windows touch without gap or overlap in both directions."""
    n_rows = len(plan) // tiles_x
    for ty in range(n_rows):
        for tx in range(tiles_x):
            t = plan[ty * tiles_x + tx]
            if tx + 1 < tiles_x and plan[ty * tiles_x + tx + 1][6] != t[7]:
                return False
            if ty + 1 < n_rows and plan[(ty + 1) * tiles_x + tx][4] != t[5]:
                return False
            if t[4] != plan[ty * tiles_x][4] or t[5] != plan[ty * tiles_x][5]:
                return False
    return True


class HostBand:
    """Host copies of the lowres rows [row0, row0 + X.shape[2]) of the four continent grids (X (1,1,h,W), W1
    (1,1,10h,10W), W2 (1,2,2h,2W), W3 (1,1,h,W)) of a grid with ``full_rows`` lowres rows: what a rank of a
    multi-GPU run keeps (and pins) on the host instead of the whole 10.9 GB continent -- ``rank_row_band`` says which
    rows its tiles read."""

    def __init__(self, X, W1, W2, W3, row0: int = 0, full_rows: Optional[int] = None):
        self.arrays = (X, W1, W2, W3)
        self.row0 = int(row0)
        self.full_rows = int(full_rows) if full_rows is not None else int(X.shape[2]) + self.row0
        h = X.shape[2]
        if tuple(W1.shape[2:]) != (10 * h, 10 * X.shape[3]) or W2.shape[2] != 2 * h or W3.shape[2] != h:
            raise ValueError("W1/W2/W3 must be 10x/2x/1x the BEDMAP2 band")


class HostDEM:
    """Pinned host output grid (1, final_y, final_x) that every rank of a run can write: one process -> a pinned
    torch tensor; torch.distributed initialised -> POSIX shared memory created by rank 0, mapped by every rank and
    registered with the CUDA driver (cudaHostRegister), so each GPU copies its finished tile rows straight into the
    final DEM over its own PCIe link (no gather, no re-placement on rank 0). ``array`` is the NumPy view."""

    def __init__(self, final_shape, dtype: str = "float32"):
        self.dtype = {"float32": torch.float32, "int16": torch.int16}[dtype]
        self.shape = (1, int(final_shape[0]), int(final_shape[1]))
        dist, rank, world = _dist()
        self._shm = None
        self._registered = False
        nbytes = self.shape[1] * self.shape[2] * (4 if dtype == "float32" else 2)
        if world == 1:
            self.tensor = torch.empty(self.shape, dtype=self.dtype, pin_memory=torch.cuda.is_available())
        else:
            from multiprocessing import resource_tracker, shared_memory
            name = [None]
            if rank == 0:
                self._shm = shared_memory.SharedMemory(create=True, size=nbytes)
                name = [self._shm.name]
            dist.broadcast_object_list(name, src=0)
            if rank != 0:
                self._shm = shared_memory.SharedMemory(name=name[0])
                try:   # Python < 3.13 registers attachments too and would unlink at this process's exit
                    resource_tracker.unregister(self._shm._name, "shared_memory")
                except Exception:
                    pass
            np_dtype = np.float32 if dtype == "float32" else np.int16
            arr = np.ndarray(self.shape, dtype=np_dtype, buffer=self._shm.buf)
            self.tensor = torch.from_numpy(arr)
            if torch.cuda.is_available():
                rc = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), nbytes, 0)
                if int(rc) != 0:
                    raise RuntimeError(f"cudaHostRegister of the shared host DEM failed ({rc})")
                self._registered = True
            dist.barrier()
        self.owner = rank == 0
        self.array = self.tensor.numpy()

    def close(self):
        dist, rank, world = _dist()
        if self._registered:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            self._registered = False
        if self._shm is not None:
            if dist is not None:
                dist.barrier()
            self.array = None
            self.tensor = None
            try:
                self._shm.close()
            except BufferError:
                pass
            if self.owner:
                self._shm.unlink()
            self._shm = None


def group_by_shape(indexed_tiles):
    groups: "OrderedDict[Tuple[int, int], List]" = OrderedDict()
    for i, t in indexed_tiles:
        groups.setdefault((t[1] - t[0], t[3] - t[2]), []).append((i, t))
    return groups


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def gather_and_assemble(results: torch.Tensor, plan, final_shape, ary_shape, rank: int, world: int,
                        new_canvas: Callable[[], torch.Tensor], place: Callable) -> Optional[torch.Tensor]:
    """Final gather: every rank contributes its (max_tiles_per_rank, ary_y, ary_x) result stack;
    rank 0 places slot ``i - start(rank_of_i)`` of each stack at the tile's canvas window.
    ``place(src_tile, canvas, ys, xs, hh, ww)`` copies src[:hh, :ww] into canvas[ys:, xs:]."""
    import torch.distributed as dist
    gathered = [torch.empty_like(results) for _ in range(world)] if rank == 0 else None
    dist.gather(results, gathered, dst=0)
    if rank != 0:
        return None
    canvas = new_canvas()
    for r in range(world):
        a, b = rank_tile_range(len(plan), r, world)
        for i in range(a, b):
            _, _, _, _, ys, ye, xs, xe = plan[i]
            place(gathered[r][i - a], canvas, ys, xs, ye - ys, xe - xs)
    return canvas


# ---- device side -----------------------------------------------------------------------------------
def _check_grids(X, W1, W2, W3):
    for a, c in ((X, 1), (W1, 1), (W2, 2), (W3, 1)):
        if a.ndim != 4 or a.shape[0] != 1 or a.shape[1] != c:
            raise ValueError(f"continent grids must be (1,C,H,W); got {tuple(a.shape)}")
    Hs, Ws = X.shape[2], X.shape[3]
    if tuple(W1.shape[2:]) != (10 * Hs, 10 * Ws) or tuple(W2.shape[2:]) != (2 * Hs, 2 * Ws) or \
            tuple(W3.shape[2:]) != (Hs, Ws):
        raise ValueError("W1/W2/W3 must be 10x/2x/1x the BEDMAP2 grid")
    return Hs, Ws


class ContinentGrids:
    """The four input rasters resident on the device: X (1,1,H,W), W1 (1,1,10H,10W), W2 (1,2,2H,2W),
    W3 (1,1,H,W); the reference's shapes are 4502x5502 etc. (deepbedmap.ipynb:1519). ``row0`` is the
    lowres row of the full grid that row 0 of these (possibly band-cropped) arrays corresponds to."""

    def __init__(self, X, W1, W2, W3, rows: Optional[Tuple[int, int]] = None):
        from .model import as_device
        Hs, _ = _check_grids(X, W1, W2, W3)
        r0, r1 = rows if rows is not None else (0, Hs)
        self.row0 = r0
        self.X = as_device(X[:, :, r0:r1])
        self.W1 = as_device(W1[:, :, 10 * r0:10 * r1])
        self.W2 = as_device(W2[:, :, 2 * r0:2 * r1])
        self.W3 = as_device(W3[:, :, r0:r1])
        self._uploaded = None  # fully resident

    @property
    def rows(self):
        return self.row0, self.row0 + self.X.shape[2]

    def ensure_rows(self, upto: int, prefetch_upto: Optional[int] = None):
        """Resident grids: nothing to do."""

    def enqueue_rows(self, upto: int, prefetch_upto: Optional[int] = None, x_first: int = 0,
                     x_last: Optional[int] = None, keep_from: Optional[int] = None):
        """Resident grids: nothing to do."""

    def wait_for(self, upto: int, x_upto: int, x_from: int = 0):
        """Resident grids: nothing to do."""


class StreamedGrids(ContinentGrids):
    """Host grids uploaded band by band on a side stream while earlier tile rows are computing
    (the host->device copy of the 10.9 GB continent overlaps the generator)."""

    def __init__(self, X, W1, W2, W3, rows: Optional[Tuple[int, int]] = None, host_row0: int = 0):
        """``host_row0``: lowres row of the full grid that row 0 of the HOST arrays corresponds to (a HostBand)."""
        from . import ops
        Hs, Ws = _check_grids(X, W1, W2, W3)
        r0, r1 = rows if rows is not None else (host_row0, host_row0 + Hs)
        if r0 < host_row0 or r1 > host_row0 + Hs:
            raise ValueError(f"host band rows [{host_row0}, {host_row0 + Hs}) do not cover the needed rows [{r0}, {r1})")
        self.row0 = r0
        self._host_row0 = host_row0
        nr = r1 - r0
        to_t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a, dtype=np.float32))
        self._host = [to_t(a) for a in (X, W1, W2, W3)]
        if any(t.dtype != torch.float32 for t in self._host):
            self._host = [t.float() for t in self._host]
        self.X = ops.empty(1, 1, nr, Ws)
        self.W1 = ops.empty(1, 1, 10 * nr, 10 * Ws)
        self.W2 = ops.empty(1, 2, 2 * nr, 2 * Ws)
        self.W3 = ops.empty(1, 1, nr, Ws)
        self._dev = [self.X, self.W1, self.W2, self.W3]
        self._scale = [1, 10, 2, 1]
        self._done = 0  # lowres rows (relative to row0) already enqueued
        self._events = []
        self._stream = torch.cuda.Stream()
        self._stream.wait_stream(torch.cuda.current_stream())

    COL_BLOCKS = 11  # a band is uploaded in column blocks (two tiles wide) so that the first tiles of a row need not wait for all of it

    def _enqueue(self, upto: int, x_first: int = 0, x_last: Optional[int] = None, keep_from: Optional[int] = None):
        """Upload of lowres rows [done, upto) in column blocks, starting with the block that holds lowres column
        ``x_first``: a rank whose tile run starts in the middle of a tile row needs that part of the band first. The
        blocks to its left follow, and only from absolute row ``keep_from`` on (the first row the NEXT tile row reads;
        None: not at all) -- nothing of this run reads the rest. Blocks wholly right of ``x_last`` (a run ending in the
        middle of its last tile row) are not uploaded."""
        from . import ops
        upto = min(upto - self.row0, self.X.shape[2])
        if upto <= self._done:
            return
        Ws = self.X.shape[3]
        edges = [Ws * k // self.COL_BLOCKS for k in range(self.COL_BLOCKS + 1)]
        blocks = list(zip(edges[:-1], edges[1:]))
        k0 = max((k for k, (xa, _) in enumerate(blocks) if xa <= x_first), default=0)
        todo = [(xa, xb, self._done) for xa, xb in blocks[k0:] if x_last is None or xa < x_last]
        if k0 > 0 and keep_from is not None:
            r0 = max(self._done, min(keep_from - self.row0, upto))
            if r0 < upto:
                todo += [(xa, xb, r0) for xa, xb in blocks[:k0]]
        elif k0 > 0 and x_first == 0:
            todo += [(xa, xb, self._done) for xa, xb in blocks[:k0]]
        pinned = all(t.is_pinned() for t in self._host)
        with torch.cuda.stream(self._stream):
            st = self._stream.cuda_stream
            for xa, xb, r0 in todo:
                for host, dev, sc in zip(self._host, self._dev, self._scale):
                    a, b = sc * r0, sc * upto
                    ha = sc * (self.row0 - self._host_row0)
                    for c in range(dev.shape[1]):
                        src, dst = host[0, c, ha + a:ha + b, sc * xa:sc * xb], dev[0, c, a:b, sc * xa:sc * xb]
                        if pinned:   # strided 2-D copy straight from the pinned grid (cudaMemcpy2DAsync)
                            ops.call("dbm_copy2d_async", dst.data_ptr(), dev.shape[3] * 4, src.data_ptr(),
                                     host.shape[3] * 4, (xb - xa) * sc * 4, b - a, st)
                        else:
                            dst.copy_(src, non_blocking=True)
                # (first row of the upload, columns of the block, event): appended in stream order
                self._events.append((r0, xa, xb, self._stream.record_event()))
        self._done = upto

    def enqueue_rows(self, upto: int, prefetch_upto: Optional[int] = None, x_first: int = 0,
                     x_last: Optional[int] = None, keep_from: Optional[int] = None):
        """Enqueue (without waiting) the upload of lowres rows < ``upto`` (see ``_enqueue`` for the column window of a
        run's first / last tile row), then of rows < ``prefetch_upto`` (one tile row ahead), on the copy stream."""
        self._enqueue(upto, x_first, x_last, keep_from)
        if prefetch_upto is not None:
            self._enqueue(prefetch_upto)

    def wait_for(self, upto: int, x_upto: int, x_from: int = 0):
        """Make the compute stream wait for rows < ``upto`` (absolute lowres row) x columns [``x_from``, ``x_upto``)
        only. The copy stream is in order, so the LAST enqueued block that holds any of those elements covers every
        earlier one."""
        need = min(upto - self.row0, self.X.shape[2])
        last = None
        for row_start, xa, xb, ev in self._events:
            if row_start < need and xa < x_upto and xb > x_from:
                last = ev
        if last is not None:
            torch.cuda.current_stream().wait_event(last)

    def ensure_rows(self, upto: int, prefetch_upto: Optional[int] = None):
        """Enqueue the upload of lowres rows up to ``prefetch_upto`` (one tile row ahead) on the copy
        stream and make the compute stream wait only for the rows < ``upto`` it is about to read."""
        self._enqueue(upto)
        self.wait_for(upto, self.X.shape[3])
        if prefetch_upto is not None:
            self._enqueue(prefetch_upto)


def predict_continent(model, X, W1=None, W2=None, W3=None, final_shape=(18000, 22000), ary_shape=(1000, 1000),
                      stride=(1000, 1000), xtrapad=(18, 18), batch_tiles: int = 4, to_host: bool = True,
                      grids: Optional[ContinentGrids] = None, out=None, out_dtype: str = "float32"):
    """Returns Y_hat (1, final_y, final_x) float32 (NumPy if ``to_host`` else a CUDA tensor), NaN
    where the reference leaves NaN. On ranks != 0 of a distributed run returns None.
    ``X`` may be a ``HostBand`` (then W1..W3 are omitted): the rows of the host grids this rank reads.
    ``out``: optional pinned host tensor (1, final_y, final_x) to receive the result without a pageable staging
    copy, or a ``HostDEM``: then every rank copies each finished tile row of its device canvas straight into that
    (shared, pinned) host grid on a copy stream while the next tiles compute, and the run ends with a barrier instead
    of a gather + re-placement + one 1.58 GB read-back on rank 0.
    ``out_dtype="int16"`` returns ``Y_hat.astype(np.int16)`` -- what the reference writes to the GeoTIFF
    (deepbedmap.py:751) -- converted on the device, which halves the device->host read."""
    if out_dtype not in ("float32", "int16"):
        raise ValueError("out_dtype must be 'float32' or 'int16'")
    from . import ops
    dist, rank, world = _dist()
    plan = tile_plan(final_shape, ary_shape, stride, xtrapad)
    tiles_x = len(range(0, final_shape[1], stride[1]))
    a, b = rank_tile_range(len(plan), rank, world)
    if grids is None:
        band = rank_row_band(plan, rank, world) if world > 1 else None
        if isinstance(X, HostBand):
            hb = X
            if band is None:
                band = (hb.row0, hb.row0 + hb.arrays[0].shape[2])
            grids = StreamedGrids(*hb.arrays, rows=band, host_row0=hb.row0)
        else:
            on_device = all(isinstance(t, torch.Tensor) and t.is_cuda for t in (X, W1, W2, W3))
            grids = (ContinentGrids if on_device else StreamedGrids)(X, W1, W2, W3, rows=band)
    g = grids
    Hs, Ws = g.X.shape[2], g.X.shape[3]
    r0, r1 = g.rows
    for (y0, y1, x0, x1, *_r) in plan[a:b]:
        if y0 < r0 or y1 > r1 or x1 > Ws:
            raise ValueError("final_shape exceeds the input grids")
    # tile rows in order (so a streamed upload can stay just ahead), same-shape tiles batched
    rows_of_tiles: "OrderedDict[int, List]" = OrderedDict()
    for i in range(a, b):
        rows_of_tiles.setdefault(i // tiles_x, []).append((i, plan[i]))
    py, px = xtrapad[0] * 4, xtrapad[1] * 4
    st = ops.stream
    stream_out = isinstance(out, HostDEM)
    if stream_out:
        if not to_host or tuple(out.shape) != (1, final_shape[0], final_shape[1]):
            raise ValueError("HostDEM output needs to_host=True and a DEM of shape (1, final_y, final_x)")
        if out.dtype != (torch.int16 if out_dtype == "int16" else torch.float32):
            raise ValueError(f"HostDEM dtype does not match out_dtype={out_dtype}")
        if not windows_abut(plan, tiles_x):
            raise ValueError("a HostDEM needs abutting tile windows (stride == ary_shape, as in the reference)")
        segs = {sg[4] // tiles_x: sg for sg in output_segments(plan, a, b, final_shape, tiles_x)}
        # local canvas: the rows this rank's segments span, full width, NaN where no tile is placed
        row_lo = min(sg[0] for sg in segs.values()) if segs else 0
        row_hi = max(sg[1] for sg in segs.values()) if segs else 0
        canvas = ops.empty(max(row_hi - row_lo, 1), final_shape[1])
        ops.fill(canvas, float("nan"))
        canvas16 = torch.empty(canvas.shape, dtype=torch.int16, device="cuda") if out_dtype == "int16" else None
        copy_stream = _copy_stream()
        results = None
    elif world == 1:
        row_lo = 0
        canvas = ops.empty(final_shape[0], final_shape[1])
        ops.fill(canvas, float("nan"))
        results = None
    else:
        results = ops.empty(max_tiles_per_rank(len(plan), world), ary_shape[0], ary_shape[1])
        ops.fill(results, float("nan"))
    row_list = list(rows_of_tiles.items())
    trace = [] if os.environ.get("DBM_TILER_TRACE") else None   # (label, host seconds, CUDA event on the compute stream)
    if trace is not None:
        import time as _time
        def mark(label):
            trace.append((label, _time.perf_counter(), torch.cuda.current_stream().record_event(
                torch.cuda.Event(enable_timing=True))))
        mark("start")
    for k, (ty, row_tiles) in enumerate(row_list):
        nxt = max(t[1] for _, t in row_list[k + 1][1]) if k + 1 < len(row_list) else None
        rows_upto = max(t[1] for _, t in row_tiles)
        # A run may start / end in the middle of a tile row: of its first row's band the part it needs goes first (the
        # rest only as far as the next tile row reads it), of its last row's band only the part it needs is uploaded.
        def band(kk):
            tl = row_list[kk][1]
            last = kk == len(row_list) - 1
            return dict(x_first=min(t[2] for _, t in tl) if kk == 0 else 0,
                        x_last=max(t[3] for _, t in tl) if last else None,
                        keep_from=min(t[0] for _, t in row_list[kk + 1][1]) if (kk == 0 and not last) else None)
        g.enqueue_rows(rows_upto, **band(k))
        if nxt is not None:
            g.enqueue_rows(nxt, **band(k + 1))
        # same-shape tiles in batches, batches ordered by their right-most grid column: a streamed upload arrives in
        # column blocks, so the left batches of a tile row start while its right part is still on the wire
        chunks = []
        for (h, w), tiles in group_by_shape(row_tiles).items():
            for b0 in range(0, len(tiles), batch_tiles):
                chunks.append(((h, w), tiles[b0:b0 + batch_tiles]))
        chunks.sort(key=lambda c: max(t[3] for _, t in c[1]))
        for (h, w), chunk in chunks:
            g.wait_for(rows_upto, max(t[3] for _, t in chunk), min(t[2] for _, t in chunk))
            nb = len(chunk)
            xb = ops.empty(nb, 1, h, w)
            w1b = ops.empty(nb, 1, 10 * h, 10 * w)
            w2b = ops.empty(nb, 2, 2 * h, 2 * w)
            w3b = ops.empty(nb, 1, h, w)
            for j, (_, (y0, y1, x0, x1, *_r)) in enumerate(chunk):
                yy = y0 - r0
                ops.call("dbm_crop_clip_f32", g.X.data_ptr(), Hs, Ws, xb[j].data_ptr(), 1, yy, x0, h, w, 0, st())
                ops.call("dbm_crop_clip_f32", g.W1.data_ptr(), 10 * Hs, 10 * Ws, w1b[j].data_ptr(), 1, 10 * yy,
                         10 * x0, 10 * h, 10 * w, 1, st())
                ops.call("dbm_crop_clip_f32", g.W2.data_ptr(), 2 * Hs, 2 * Ws, w2b[j].data_ptr(), 2, 2 * yy, 2 * x0,
                         2 * h, 2 * w, 1, st())
                ops.call("dbm_crop_clip_f32", g.W3.data_ptr(), Hs, Ws, w3b[j].data_ptr(), 1, yy, x0, h, w, 1, st())
            y = model.forward(xb, w1b, w2b, w3b).array  # (nb, 1, 4(h-2), 4(w-2))
            th, tw = y.shape[2], y.shape[3]
            for j, (i, (y0, y1, x0, x1, ys, ye, xs, xe)) in enumerate(chunk):
                hh, ww = ye - ys, xe - xs
                # the reference assigns Y_pred[72:-72, 72:-72] into Y_hat[ys:ye, xs:xe] and raises
                # ValueError on a shape mismatch (deepbedmap.py:734-738)
                if (th - 2 * py, tw - 2 * px) != (hh, ww):
                    raise ValueError(f"could not broadcast tile {(th - 2 * py, tw - 2 * px)} into {(hh, ww)}")
                if results is None:
                    ops.call("dbm_place_tile_f32", y[j].data_ptr(), th, tw, py, px, canvas.data_ptr(),
                             canvas.shape[0], final_shape[1], ys - row_lo, xs, hh, ww, st())
                else:
                    ops.call("dbm_place_tile_f32", y[j].data_ptr(), th, tw, py, px, results[i - a].data_ptr(),
                             ary_shape[0], ary_shape[1], 0, 0, hh, ww, st())
            del y, xb, w1b, w2b, w3b
            if trace is not None:
                mark(f"row{ty} batch x<={max(t[3] for _, t in chunk)}")
            if stream_out and canvas16 is None:
                # these tiles' rectangles of the local canvas are final: hand them to the copy stream now (per batch, not
                # per tile row: the device->host stream of a row overlaps the rest of that row, and what is left
                # after the last batch of a run is one batch, not 88 MB)
                sg = segs[ty]
                copy_stream.wait_stream(torch.cuda.current_stream())
                pitch = final_shape[1] * 4
                for i, t in chunk:
                    xs_t = sg[2] if i == sg[4] else t[6]
                    xe_t = sg[3] if i == sg[5] - 1 else t[7]
                    ops.call("dbm_copy2d_async", out.tensor.data_ptr() + sg[0] * pitch + xs_t * 4, pitch,
                             canvas.data_ptr() + (sg[0] - row_lo) * pitch + xs_t * 4, pitch, (xe_t - xs_t) * 4,
                             sg[1] - sg[0], copy_stream.cuda_stream)
        if stream_out and canvas16 is not None:
            # int16 product: this tile row of the local canvas is final -- convert it and hand its rectangle to the copy
            # stream (the float32 product went out batch by batch above)
            ys, ye, xs, xe, _, _ = segs[ty]
            src, esz = canvas, 4
            if canvas16 is not None:
                o = (ys - row_lo) * final_shape[1]
                ops.call("dbm_f32_to_i16", canvas.data_ptr() + 4 * o, canvas16.data_ptr() + 2 * o,
                         (ye - ys) * final_shape[1], st())
                src, esz = canvas16, 2
            copy_stream.wait_stream(torch.cuda.current_stream())
            pitch = final_shape[1] * esz
            ops.call("dbm_copy2d_async", out.tensor.data_ptr() + ys * pitch + xs * esz, pitch,
                     src.data_ptr() + (ys - row_lo) * pitch + xs * esz, pitch, (xe - xs) * esz, ye - ys,
                     copy_stream.cuda_stream)
    if stream_out:
        if trace is not None:
            mark("enqueued")
        copy_stream.synchronize()
        if trace is not None:
            t_copy = _time.perf_counter()
        if dist is not None:
            dist.barrier()
        if trace is not None:
            t_end = _time.perf_counter()
            torch.cuda.synchronize()
            h0, e0 = trace[0][1], trace[0][2]
            print(f"[tiler trace rank {rank}] host: all enqueued {1e3 * (trace[-1][1] - h0):.1f} ms, copies done "
                  f"{1e3 * (t_copy - h0):.1f} ms, barrier done {1e3 * (t_end - h0):.1f} ms; device (compute stream): "
                  + ", ".join(f"{lab} {e0.elapsed_time(ev):.1f}" for lab, _, ev in trace[1:]), flush=True)
        return out.array if rank == 0 else None
    if world > 1:
        def new_canvas():
            c = ops.empty(final_shape[0], final_shape[1])
            ops.fill(c, float("nan"))
            return c

        def place(src, canvas_, ys, xs, hh, ww):
            ops.call("dbm_place_tile_f32", src.data_ptr(), ary_shape[0], ary_shape[1], 0, 0, canvas_.data_ptr(),
                     final_shape[0], final_shape[1], ys, xs, hh, ww, st())

        canvas = gather_and_assemble(results, plan, final_shape, ary_shape, rank, world, new_canvas, place)
        if canvas is None:
            return None
    res = canvas.view(1, final_shape[0], final_shape[1])
    if out_dtype == "int16":
        res16 = torch.empty(res.shape, dtype=torch.int16, device=res.device)
        ops.call("dbm_f32_to_i16", res.data_ptr(), res16.data_ptr(), res.numel(), st())
        res = res16
    if not to_host:
        return res
    if out is not None:
        if tuple(out.shape) != tuple(res.shape) or out.dtype != res.dtype or out.is_cuda:
            raise ValueError(f"out must be a host {out_dtype} tensor of shape (1, final_y, final_x)")
        out.copy_(res, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out.numpy()
    return res.cpu().numpy()


_COPY_STREAM = None


def _copy_stream():
    global _COPY_STREAM
    if _COPY_STREAM is None:
        _COPY_STREAM = torch.cuda.Stream()
    return _COPY_STREAM
