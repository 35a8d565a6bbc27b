"""Prints cycles per tcgen05.mma (M=128, K=16) for the operand layouts / N considered."""
import ctypes, sys
import torch
sys.path.insert(0, ".")
import os
from deepbedmap_b200 import _lib, build
_lib.load()
lib = ctypes.CDLL(build.TUNING_LIB)   # the microbenchmark lives in its own library, not in the product one
lib.dbm_last_error = _lib.load().dbm_last_error
lib.dbm_debug_umma_rate.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
lib.dbm_debug_umma_rate.restype = ctypes.c_int
out = torch.zeros(148, dtype=torch.int64, device="cuda")
for grid in (1, 148):
    for n in (32, 64, 128, 256):
        for mode, name in ((0, "noswz-halo"), (1, "noswz-dense"), (2, "sw128")):
            iters = 400
            rc = lib.dbm_debug_umma_rate(mode, n, iters, 36, out.data_ptr(), grid, None)
            torch.cuda.synchronize()
            assert rc == 0, lib.dbm_last_error()
            cyc = out[:grid].double().mean().item() / (iters * 36)
            smem_b = 128 * 32 + n * 32
            print(f"grid={grid:3d} N={n:3d} {name:12s}: {cyc:7.1f} cyc/MMA (tensor floor {n/2:.0f}) -> "
                  f"{n/2/cyc*100:5.1f}% of tensor peak, operand bytes/cyc {smem_b/cyc:6.1f}")
