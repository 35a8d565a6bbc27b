"""A/B of the trunk kernel's pass plans on interior continent tiles: both dense-block pairs fused (default), only the
second pair (conv3 + conv4) fused, no pairing; for batches of 1, 2 and 4 tiles. Prints ms and algorithmic TFLOP/s."""
import sys
import torch
sys.path.insert(0, ".")
from deepbedmap_b200 import GeneratorModel

H = W = 286
m = GeneratorModel(precision="bf16")
for n in (1, 2, 4):
    for name, paired, convs in (("both pairs", True, (1, 3)), ("second pair", True, (3,)), ("unpaired", False, (1, 3))):
        m.paired_trunk, m.paired_convs = paired, convs
        ws = m._trunk_workspace(n, H, W)
        ws["s0"].normal_()
        for _ in range(3):
            m._run_trunk(ws, n, H, W)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 40
        e0.record()
        for _ in range(reps):
            m._run_trunk(ws, n, H, W)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"n={n} {name:12s}: {ms:7.3f} ms  {ws['flops'] / ms / 1e9:7.1f} TFLOP/s", flush=True)
