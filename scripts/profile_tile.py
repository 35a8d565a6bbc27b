"""Runs the bf16 generator forward on a batch of interior continent tiles (288x288 lowres ->
1144x1144); used under ncu to get the per-kernel launch list / full captures."""
import sys
import torch
sys.path.insert(0, ".")
from deepbedmap_b200 import GeneratorModel

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
precision = sys.argv[3] if len(sys.argv) > 3 else "bf16"   # bf16 | bf16x3 | fp32
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(n, 1, 288, 288, generator=g, device="cuda") * 800 - 500
w1 = torch.rand(n, 1, 2880, 2880, generator=g, device="cuda") * 4000
w2 = torch.randn(n, 2, 576, 576, generator=g, device="cuda").clamp_(min=0) * 200
w3 = torch.rand(n, 1, 288, 288, generator=g, device="cuda") * 1000
m = GeneratorModel(precision=precision)
for _ in range(reps):
    y = m.forward(x, w1, w2, w3).array
torch.cuda.synchronize()
print("ok", tuple(y.shape), float(y.abs().mean()))
