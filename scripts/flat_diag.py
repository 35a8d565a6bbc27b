"""Diagnostics of the tensor-core training trunk on a GPU box: MN-major descriptor convention check
(wgrad with LBO/SBO as documented vs swapped) and CUDA-event timings of the trunk's forward chain,
data-gradient chain and batched weight gradient at the training shape (batch 128, 9x9 trunk pixels)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepbedmap_b200 import GeneratorModel, flat, ops  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def wgrad_once(swap, n=3, h=9, w=9, cin=160):
    ops.call("dbm_flat_debug_set", 1, swap)
    geom = flat.geometry(n, h, w)
    pg = geom["Pg"]
    g0 = torch.Generator().manual_seed(1)
    a = torch.randn(n, cin, h, w, generator=g0).cuda()
    g = torch.randn(n, 32, h, w, generator=g0).cuda()
    ab, gb = flat.alloc_bf16(cin, geom), flat.alloc_bf16(32, geom)
    flat.from_nchw(a, dst8=ab)
    flat.from_nchw(g, dst8=gb)
    dw = ops.zeros(32, cin, 3, 3)
    units, reduces = [], []
    for c0, nch in flat.chunk_channels(cin):
        first = len(units)
        for blk0, nblk in flat.split_blocks(geom["tiles"], 2):
            units.append((ab.data_ptr() + 2 * c0 * pg, gb.data_ptr(), len(units), blk0, nblk, nch // 8, (0, 0, 0)))
        reduces.append((first, dw.data_ptr(), flat.PARTIAL_FLOATS, 2, cin, c0, 0, nch, 0))
    partial = torch.zeros(len(units) * flat.PARTIAL_FLOATS, device="cuda")
    u = np.array(units, dtype=flat.WGRAD_UNIT_DTYPE)
    u["partial"] = partial.data_ptr() + u["partial"] * np.uint64(flat.PARTIAL_FLOATS * 4)
    rd = np.array(reduces, dtype=flat.WGRAD_REDUCE_DTYPE)
    rd["partial"] = partial.data_ptr() + rd["partial"] * np.uint64(flat.PARTIAL_FLOATS * 4)
    dev = lambda t: torch.from_numpy(t.view(np.uint8).reshape(-1).copy()).cuda()
    ud, rdd = dev(u), dev(rd)
    ops.call("dbm_flat_wgrad", ud.data_ptr(), len(u), n, h, w, ops.stream())
    ops.call("dbm_flat_wgrad_reduce", rdd.data_ptr(), len(rd), ops.stream())
    torch.cuda.synchronize()
    bfr = lambda t: t.to(torch.bfloat16).double().cpu()
    wz = torch.zeros(32, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(bfr(a), wz, padding=1).backward(bfr(g))
    ops.call("dbm_flat_debug_set", 1, 0)
    return rel(dw, wz.grad)


def timings(nb=12, n=128):
    m = GeneratorModel(num_residual_blocks=nb)
    ft = m._flat_trunk(n, 9, 9)
    a0 = torch.randn(n, 128, 9, 9, device="cuda")
    da3 = torch.randn(n, 64, 9, 9, device="cuda")
    st = ops.stream()
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    t_f = timed(lambda: ft.forward(a0))
    t_d = timed(lambda: ops.call("dbm_flat_conv3x3_seq", ft.bwd.ctypes.data, len(ft.bwd), n, 9, 9, 0, 0, st))
    t_w = timed(lambda: ops.call("dbm_flat_wgrad", ft.units_dev.data_ptr(), ft.n_units, n, 9, 9, st))
    t_r = timed(lambda: (ops.call("dbm_flat_wgrad_reduce", ft.reduce_dev.data_ptr(), ft.n_reduce, st),
                         ops.call("dbm_flat_bias_grad", ft.bias_dev.data_ptr(), ft.n_bias, n, 9, 9, st)))
    t_b = timed(lambda: ft.backward(da3))
    fl = ft.flops_fwd
    print(f"flat trunk nb={nb} n={n}: forward {t_f:.3f} ms ({len(ft.fwd)} launches, {fl / t_f / 1e9:.1f} TFLOP/s), "
          f"dgrad chain {t_d:.3f} ms ({len(ft.bwd)} launches, {fl / t_d / 1e9:.1f} TFLOP/s), wgrad {t_w:.3f} ms "
          f"({ft.n_units} units, {fl / t_w / 1e9:.1f} TFLOP/s), reduce+bias {t_r:.3f} ms, backward total {t_b:.3f} ms")


if __name__ == "__main__":
    torch.cuda.set_device(0)
    for swap in (0, 1):
        try:
            print(f"wgrad MN-major descriptors swap={swap}: rel_l2 vs fp64 autograd = {wgrad_once(swap):.3e}", flush=True)
        except Exception as ex:  # a trap poisons the context: report and stop
            print(f"wgrad swap={swap}: FAILED {ex!r}", flush=True)
            break
    try:
        timings()
    except Exception as ex:
        print(f"timings FAILED {ex!r}")
