"""Times the training trunk forward (batch of 9x9 tiles): image-resident kernel vs the flat chain."""
import sys
import torch
sys.path.insert(0, ".")
from deepbedmap_b200 import GeneratorModel

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
m = GeneratorModel(num_residual_blocks=12)
ft = m._flat_trunk(n, 9, 9)
a0 = torch.randn(n, 128, 9, 9, device="cuda")
for local in (False, True):
    ft.local = local
    for _ in range(3):
        ft.forward(a0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ft.forward(a0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"n={n} local={local}: {ms:.3f} ms per trunk forward (incl. layout conversions), {ft.flops_fwd / ms / 1e9:.1f} TFLOP/s")
