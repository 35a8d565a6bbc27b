"""BASELINE.json configs[1] and configs[4]: GeneratorModel forward on the
tensor-core path over the reference's hyperparameter search space -- num_residual_blocks
(srgan_train.py:454; 8-14 in the paper, 23 = ESRGAN's depth) x inter_channels 32 / 64 (:283-284) --
at batch 128 / 1024 training tiles (9x9 trunk px each) and on one interior continent tile (286x286).
Times the whole forward and the trunk kernel alone with CUDA events and prints algorithmic TFLOP/s
and the fraction of the measured sustained bf16 peak (MEASURED_PEAKS.json).

usage: python scripts/sweep_config5.py [quick]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepbedmap_b200 import GeneratorModel, flat, ops  # noqa: E402


def macs_per_trunk_px(nb, g):
    """Algorithmic multiply-accumulates of the generator per trunk (lowres) pixel."""
    stem = 32 * (9 + 900 + 72 + 9)
    rdb = 9 * (64 * g + (64 + g) * g + (64 + 2 * g) * g + (64 + 3 * g) * g + (64 + 4 * g) * 64)
    trunk = 9 * 128 * 64 + 3 * nb * rdb + 9 * 64 * 64
    head = 4 * 9 * 64 * 64 + 16 * 9 * 64 * 64 + 16 * 9 * 64 * (18 + 64) + 16 * 9 * 64 * (18 + 1)
    return stem + trunk + head, trunk


def timed(fn, warm=3, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk)).get("bf16_tflops_sustained", 1383.0) if os.path.exists(pk) else 1400.0
    depths = (8, 12, 23) if quick else (8, 10, 12, 14, 23)
    shapes = [("128 tiles 11x11", 128, 11, 11), ("1024 tiles 11x11", 1024, 11, 11), ("1 tile 288x288", 1, 288, 288)]
    print(f"bf16 sustained peak {peak:.1f} TFLOP/s (MEASURED_PEAKS.json); activations exceed L2 at every size below "
          f"except 128 tiles; forward = whole generator, trunk = the trunk kernel forward() runs, alone: local_trunk_kernel on 11x11 tiles with inter_channels 32, else umma_trunk_kernel (CUDA events)")
    print(f"{'workload':18s} {'nb':>3s} {'inter':>5s} {'GFLOP fwd':>10s} {'fwd ms':>8s} {'fwd TF/s':>9s} {'frac':>6s} "
          f"{'trunk ms':>9s} {'trunk TF/s':>10s} {'frac':>6s}")
    for label, n, h, w in shapes:
        gen = torch.Generator(device="cuda").manual_seed(42)
        ins = (torch.rand(n, 1, h, w, generator=gen, device="cuda"), torch.rand(n, 1, 10 * h, 10 * w, generator=gen, device="cuda"),
               torch.rand(n, 2, 2 * h, 2 * w, generator=gen, device="cuda"), torch.rand(n, 1, h, w, generator=gen, device="cuda"))
        for inter in (32, 64):
            for nb in depths:
                m = GeneratorModel(num_residual_blocks=nb, inter_channels=inter, precision="bf16")
                total, trunk = macs_per_trunk_px(nb, inter)
                px = n * (h - 2) * (w - 2)
                ms = timed(lambda: m.forward(*ins))
                if m.local_trunk and inter == 32 and flat.local_trunk_fits(h - 2, w - 2):
                    # small tiles: the image-resident kernel (csrc/umma_local.cu) is the trunk forward() runs
                    ws = m._local_workspace(n, h - 2, w - 2, m._pack(m.PACK_INFER_LOCAL))
                    tms = timed(lambda: ops.call("dbm_trunk_local_fwd", ws["table"].data_ptr(), ws["count"], n, h - 2, w - 2,
                                                 ws["s0"].data_ptr(), ws["x"][0].data_ptr(), ws["x"][1].data_ptr(),
                                                 ops.stream()))
                else:
                    ws = m._trunk_workspace(n, h - 2, w - 2)
                    assert abs(ws["flops"] - 2.0 * trunk * px) < 1e-6 * ws["flops"]
                    tms = timed(lambda: m._run_trunk(ws, n, h - 2, w - 2))
                tf, ttf = 2.0 * total * px / ms / 1e9, 2.0 * trunk * px / tms / 1e9
                print(f"{label:18s} {nb:3d} {inter:5d} {2.0 * total * px / 1e9:10.2f} {ms:8.3f} {tf:9.1f} {tf / peak:6.3f} "
                      f"{tms:9.3f} {ttf:10.1f} {ttf / peak:6.3f}", flush=True)
                del m, ws
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
