"""Layer-by-layer comparison of DiscriminatorModel(precision="bf16") with the bf16-operand oracle (debug aid)."""
import os, sys
import numpy as np
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepbedmap_b200 import DiscriminatorModel  # noqa: E402
from oracle import deepbedmap_oracle as O  # noqa: E402

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))

for n in (6, 64):
    params = O.init_discriminator_params(seed=1, bias_std=0.1, scale=1.0)
    outs = {}
    for prec in ("bf16", "fp32"):
        d = DiscriminatorModel(precision=prec)
        for k in d.p:
            d.set_param(k, params[k])
        x = np.random.RandomState(0).rand(n, 1, 36, 36).astype(np.float32)
        out = d.forward(x, train=True, save=True).array
        outs[prec] = (d._ctx["pres"], d._ctx["acts"], out)
    p = O.to_torch(params)
    q = O._q
    for emu in (True, False):
        qq = q if emu else (lambda t: t)
        a = F.leaky_relu(F.conv2d(torch.as_tensor(x, dtype=torch.float64), p["conv_layer0/W"], p["conv_layer0/b"], padding=1), 0.2)
        line = []
        for i in range(1, 10):
            _, k, s = O.DISC_CONVS[i]
            z = F.conv2d(qq(a), qq(p[f"conv_layer{i}/W"]), None, stride=s, padding=1)
            line.append(f"z{i} {rel(outs['bf16'][0][i - 1], z):.1e}/{rel(outs['fp32'][0][i - 1], z):.1e}")
            a = F.leaky_relu(O.batch_norm(p, f"batch_norm{i}", z, True, None), 0.2)
        print(f"n={n} emulate={emu} (bf16 model / fp32 model): " + " ".join(line))
