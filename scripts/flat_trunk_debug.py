"""Intermediate-by-intermediate comparison of the tensor-core training trunk against fp64 autograd
(debugging aid: prints the relative L2 error of every activation slot, every gradient slot, every
parameter gradient of a small trunk)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepbedmap_b200 import GeneratorModel, flat, layout  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
beta = 0.1
H = W = 9


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


m = GeneratorModel(num_residual_blocks=nb, residual_scaling=beta, seed=3, init_scale=1.0)
rs = np.random.RandomState(5)
for k in m.p:
    if k.endswith("/b"):
        m.set_param(k, rs.randn(*m.p[k].shape).astype(np.float32) * 0.05)
params = {k: m.get_param(k) for k in m.p}
p64 = {k: torch.as_tensor(v, dtype=torch.float64).requires_grad_(True) for k, v in params.items()}
g = torch.Generator().manual_seed(11)
a0 = torch.randn(n, 128, H, W, generator=g).cuda()
da3 = torch.randn(n, 64, H, W, generator=g).cuda()

ft = m._flat_trunk(n, H, W)
a3 = ft.forward(a0)
m.cleargrads()
da0 = ft.backward(da3)
torch.cuda.synchronize()

conv = lambda x, key: F.conv2d(x, p64[key + "/W"], p64[key + "/b"], padding=1)
a0r = a0.double().cpu().requires_grad_(True)
z_pre = conv(a0r, "pre_residual_conv_layer"); z_pre.retain_grad()
a1 = F.leaky_relu(z_pre, 0.2)
xs, zs = [a1], {}
cur = a1
for i in range(nb):
    rrdb_in = cur
    for r in (1, 2, 3):
        pre = f"residual_network/{i}/residual_dense_block{r}"
        feats = [cur]
        j = 3 * i + r - 1
        for k in (1, 2, 3, 4):
            z = conv(torch.cat(feats, 1), f"{pre}/conv_layer{k}"); z.retain_grad(); zs[(j, k)] = z
            feats.append(F.leaky_relu(z, 0.2))
        z5 = conv(torch.cat(feats, 1), f"{pre}/conv_layer5"); z5.retain_grad(); zs[(j, 5)] = z5
        cur = cur + beta * z5
        if r == 3:
            cur = rrdb_in + beta * cur
        cur.retain_grad()
        xs.append(cur)
z_post = conv(cur, "post_residual_conv_layer"); z_post.retain_grad()
a3r = a1 + z_post
(a3r * da3.double().cpu()).sum().backward()

print(f"forward a3 {rel(a3, a3r.detach()):.2e}   da0 {rel(da0, a0r.grad):.2e}")
nrdb = 3 * nb
for j in range(nrdb + 1):
    print(f"cat[{j}][0:64] (x_{j}) {rel(flat.to_nchw(ft.cat[j], 64, n, H, W), xs[j].detach()):.2e}", end="  ")
    if j < nrdb:
        for k in (1, 2, 3, 4):
            print(f"a{k} {rel(flat.to_nchw(ft.cat[j], 32, n, H, W, c0=32 + 32 * k), F.leaky_relu(zs[(j, k)].detach(), 0.2)):.2e}", end=" ")
    print()
print(f"gpost {rel(flat.to_nchw(ft.gpost, 64, n, H, W), z_post.grad):.2e}   gpre {rel(flat.to_nchw(ft.gpre, 64, n, H, W), z_pre.grad):.2e}")
for j in reversed(range(nrdb)):
    s = f"gcat[{j}]: g5 {rel(flat.to_nchw(ft.gcat[j], 64, n, H, W, c0=128), zs[(j, 5)].grad):.2e}"
    for k in (4, 3, 2, 1):
        s += f"  g{k} {rel(flat.to_nchw(ft.gcat[j], 32, n, H, W, c0=32 * (k - 1)), zs[(j, k)].grad):.2e}"
    print(s)
for k in m.p:
    if k.startswith(("residual_network", "pre_residual", "post_residual")):
        print(f"grad {k}: {rel(m.g[k], p64[k].grad):.2e}")
