"""GeneratorModel forward on a batch of 11x11 windows (BASELINE configs[1]: batch 128; larger batches for the
image-resident trunk kernel's throughput): run under ncu."""
import sys
import torch
sys.path.insert(0, ".")
from deepbedmap_b200 import GeneratorModel

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g = torch.Generator(device="cuda").manual_seed(0)
ins = (torch.rand(n, 1, 11, 11, generator=g, device="cuda"), torch.rand(n, 1, 110, 110, generator=g, device="cuda"),
       torch.rand(n, 2, 22, 22, generator=g, device="cuda"), torch.rand(n, 1, 11, 11, generator=g, device="cuda"))
m = GeneratorModel(precision="bf16")
for _ in range(reps):
    y = m.forward(*ins).array
torch.cuda.synchronize()
print("ok", tuple(y.shape), float(y.abs().mean()))
