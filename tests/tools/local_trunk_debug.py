"""Where does the image-resident trunk first differ from the flat chain? (per dense-block buffer)"""
import sys
import torch
sys.path.insert(0, ".")
from deepbedmap_b200 import GeneratorModel
from oracle import deepbedmap_oracle as O

for nb, n in ((2, 3), (1, 128), (3, 128), (12, 4)):
    params = O.init_generator_params(nb, seed=5, bias_std=0.05, scale=0.7)
    m = GeneratorModel(num_residual_blocks=nb)
    for k, v in params.items():
        m.set_param(k, v)
    g = torch.Generator().manual_seed(21)
    a0 = torch.randn(n, 128, 9, 9, generator=g).cuda()
    ft = m._flat_trunk(n, 9, 9)
    out = {}
    for local in (False, True):
        ft.local = local
        for c in ft.cat:
            c.zero_()
        a3 = ft.forward(a0).clone()
        out[local] = ([c.clone() for c in ft.cat], a3)
    line = []
    for j, (a, b) in enumerate(zip(out[True][0], out[False][0])):
        d = (a != b)
        if d.any():
            # which 32-channel block (slab // 4) differs first
            blocks = [int(d[4 * k:4 * k + 4].any()) for k in range(6)]
            line.append(f"j={j}:{float(d.float().mean()):.1e}{blocks}")
    print(f"nb={nb} n={n}: a3 equal {torch.equal(out[True][1], out[False][1])}; differing buffers: {line[:6]}")
