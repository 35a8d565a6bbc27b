"""
CPU oracle for the DeepBedMap ESRGAN hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch restatement (torch on CPU, float64 by default) of the
arithmetic the reference executes through Chainer 7.0.0 for its generator,
discriminator, losses, Adam update, training step and continent tiler.  It is the
parity checker and the timed host-CPU baseline; nothing under ``deepbedmap_b200/``
may import it.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it.

Pinning status
--------------
Chainer / CuPy / ssim-chainer are third-party dependencies whose sources are not under
/root/reference (Pipfile:8,10,30) and none of them is installable in this image, so the
reference as a whole cannot be executed here.  What pins this file:
  * the reference's own test values: output shapes and parameter counts (srgan_train.py:444-447,
    605-608) and the four loss / metric known answers (srgan_train.py:868, 920, 948, 991)
    -- tests/test_oracle_kat.py;
  * the reference's own CODE, executed: srgan_train.py's model classes, loss functions and both
    step functions and deepbedmap.py's tiler cell, compiled unmodified from /root/reference and run
    on a float64 stand-in for the Chainer calls they make (tests/tools/minichainer.py,
    tests/golden/make_reference_golden.py -> tests/golden/reference_graph_golden.npz).  This
    file reproduces those vectors to 1e-10 (forward, losses) / 1e-8 (parameter gradients of a
    D-step + G-step) and the tile geometry exactly -- tests/test_reference_graph.py.
So the graph, the step logic and the key layout are pinned to the reference by execution.  The
arithmetic of the Chainer PRIMITIVES under them (convolution, deformable sampler,
BatchNormalization, Adam) is restated both here and in the stand-in; no
Chainer-computed value of a conv / RRDB / deformable / BN layer exists to compare with, so at the
primitive level:  **parity unpinned**  (cross-checks: torch functional ops, a naive NumPy loop and
torchvision.ops.deform_conv2d for the deformable convolution).

Every function cites the reference lines it follows.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

LRELU_SLOPE = 0.2  # srgan_train.py:340 (F.leaky_relu slope=0.2)


# --------------------------------------------------------------------------------------
# Parameter inventory (Chainer .npz key layout)
# --------------------------------------------------------------------------------------
def generator_param_shapes(num_residual_blocks: int = 12, inter_channels: int = 32,
                           out_channels: int = 1) -> "OrderedDict[str, tuple]":
    """Keys/shapes of GeneratorModel's parameters in Chainer's serializer naming.

    Follows srgan_train.py:218-254 (input block), :281-331 (dense block), :376-391
    (RRDB), :450-523 (generator).  Order = Chainer's sorted-children traversal is not
    needed anywhere; we keep definition order.
    """
    g = inter_channels
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["input_block/conv_on_X/W"] = (32, 1, 3, 3)
    s["input_block/conv_on_X/b"] = (32,)
    s["input_block/conv_on_W1/W"] = (32, 1, 30, 30)
    s["input_block/conv_on_W1/b"] = (32,)
    s["input_block/conv_on_W2/W"] = (32, 2, 6, 6)
    s["input_block/conv_on_W2/b"] = (32,)
    s["input_block/conv_on_W3/W"] = (32, 1, 3, 3)
    s["input_block/conv_on_W3/b"] = (32,)
    s["pre_residual_conv_layer/W"] = (64, 128, 3, 3)
    s["pre_residual_conv_layer/b"] = (64,)
    for i in range(num_residual_blocks):
        for r in (1, 2, 3):
            p = f"residual_network/{i}/residual_dense_block{r}"
            for k in (1, 2, 3, 4):
                s[f"{p}/conv_layer{k}/W"] = (g, 64 + (k - 1) * g, 3, 3)
                s[f"{p}/conv_layer{k}/b"] = (g,)
            s[f"{p}/conv_layer5/W"] = (64, 64 + 4 * g, 3, 3)
            s[f"{p}/conv_layer5/b"] = (64,)
    for name in ("post_residual_conv_layer", "post_upsample_conv_layer_1",
                 "post_upsample_conv_layer_2"):
        s[f"{name}/W"] = (64, 64, 3, 3)
        s[f"{name}/b"] = (64,)
    for name, oc in (("final_conv_layer1", 64), ("final_conv_layer2", out_channels)):
        s[f"{name}/offset_conv/W"] = (18, 64, 3, 3)
        s[f"{name}/offset_conv/b"] = (18,)
        s[f"{name}/deform_conv/W"] = (oc, 64, 3, 3)
        s[f"{name}/deform_conv/b"] = (oc,)
    return s


# (out_channels, ksize, stride) of conv_layer0..9, srgan_train.py:617-634
DISC_CONVS = [(64, 3, 1), (64, 4, 2), (128, 3, 1), (128, 4, 2), (128, 3, 1),
              (256, 4, 2), (256, 3, 1), (512, 4, 2), (512, 3, 1), (512, 4, 2)]


def discriminator_param_shapes() -> "OrderedDict[str, tuple]":
    """Trainable parameters of DiscriminatorModel (srgan_train.py:611-647)."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    cin = 1
    for i, (cout, k, _) in enumerate(DISC_CONVS):
        s[f"conv_layer{i}/W"] = (cout, cin, k, k)
        if i == 0:
            s["conv_layer0/b"] = (cout,)  # only the first conv has a bias (:623)
        else:
            s[f"batch_norm{i}/gamma"] = (cout,)
            s[f"batch_norm{i}/beta"] = (cout,)
        cin = cout
    s["linear_1/W"] = (100, 512)
    s["linear_1/b"] = (100,)
    s["linear_2/W"] = (1, 100)
    s["linear_2/b"] = (1,)
    return s


def discriminator_persistent_shapes() -> "OrderedDict[str, tuple]":
    """BatchNormalization persistents serialised next to the params."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    for i, (cout, _, _) in enumerate(DISC_CONVS):
        if i:
            s[f"batch_norm{i}/avg_mean"] = (cout,)
            s[f"batch_norm{i}/avg_var"] = (cout,)
            s[f"batch_norm{i}/N"] = ()
    return s


def count_params(shapes) -> int:
    return int(sum(int(np.prod(v)) for v in shapes.values()))


def _he_normal(rng: np.random.RandomState, shape, scale=0.1) -> np.ndarray:
    """chainer.initializers.HeNormal(scale=0.1, fan_option='fan_in') (srgan_train.py:220)."""
    fan_in = int(np.prod(shape[1:]))
    std = scale * math.sqrt(2.0 / fan_in)
    return rng.normal(0.0, std, size=shape).astype(np.float32)


def init_generator_params(num_residual_blocks=12, seed=0, inter_channels=32,
                          bias_std: float = 0.0, scale: float = 0.1) -> "OrderedDict[str, np.ndarray]":
    """Seeded synthetic weights in App. C key order.  ``bias_std`` > 0 draws non-zero
    biases (Chainer's default is zeros) so that parity tests exercise the bias path;
    ``scale`` = 0.1 is the reference's HeNormal scale (srgan_train.py:220), larger values
    give trained-like O(1) activations for non-degenerate parity checks."""
    rng = np.random.RandomState(seed)
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for k, shp in generator_param_shapes(num_residual_blocks, inter_channels).items():
        if k.endswith("/W"):
            out[k] = _he_normal(rng, shp, scale)
        else:
            out[k] = (rng.normal(0, bias_std, size=shp) if bias_std else np.zeros(shp)).astype(np.float32)
    return out


def init_discriminator_params(seed=1, bias_std: float = 0.0, scale: float = 0.1) -> "OrderedDict[str, np.ndarray]":
    rng = np.random.RandomState(seed)
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for k, shp in discriminator_param_shapes().items():
        if k.endswith("/W"):
            out[k] = _he_normal(rng, shp, scale)
        elif k.endswith("/gamma"):
            out[k] = (1.0 + (rng.normal(0, bias_std, size=shp) if bias_std else 0.0)) * np.ones(shp, np.float32)
            out[k] = out[k].astype(np.float32)
        else:
            out[k] = (rng.normal(0, bias_std, size=shp) if bias_std else np.zeros(shp)).astype(np.float32)
    for k, shp in discriminator_persistent_shapes().items():
        if k.endswith("avg_mean"):
            out[k] = np.zeros(shp, np.float32)
        elif k.endswith("avg_var"):
            out[k] = np.ones(shp, np.float32)
        else:
            out[k] = np.array(0, dtype=np.int64)
    return out


def to_torch(params, dtype=torch.float64, requires_grad=False) -> Params:
    out = {}
    for k, v in params.items():
        t = torch.as_tensor(np.asarray(v))
        if t.is_floating_point():
            t = t.to(dtype).clone()
            if requires_grad and not (k.endswith("avg_mean") or k.endswith("avg_var")):
                t.requires_grad_(True)
        out[k] = t
    return out


# --------------------------------------------------------------------------------------
# Generator (srgan_train.py:201-576)
# --------------------------------------------------------------------------------------
def _lrelu(x):
    return F.leaky_relu(x, LRELU_SLOPE)


def input_block(p: Params, x, w1, w2, w3):
    """DeepbedmapInputBlock.forward, srgan_train.py:256-266 (valid-padded strided convs)."""
    x_ = F.conv2d(x, p["input_block/conv_on_X/W"], p["input_block/conv_on_X/b"], stride=1)
    w1_ = F.conv2d(w1, p["input_block/conv_on_W1/W"], p["input_block/conv_on_W1/b"], stride=10)
    w2_ = F.conv2d(w2, p["input_block/conv_on_W2/W"], p["input_block/conv_on_W2/b"], stride=2)
    w3_ = F.conv2d(w3, p["input_block/conv_on_W3/W"], p["input_block/conv_on_W3/b"], stride=1)
    return torch.cat((x_, w1_, w2_, w3_), dim=1)


def residual_dense_block(p: Params, prefix: str, x, beta: float):
    """ResidualDenseBlock.forward, srgan_train.py:333-360."""
    a0 = x
    feats = [a0]
    for k in (1, 2, 3, 4):
        a = F.conv2d(torch.cat(feats, dim=1), p[f"{prefix}/conv_layer{k}/W"],
                     p[f"{prefix}/conv_layer{k}/b"], padding=1)
        feats.append(_lrelu(a))
    a5 = F.conv2d(torch.cat(feats, dim=1), p[f"{prefix}/conv_layer5/W"],
                  p[f"{prefix}/conv_layer5/b"], padding=1)
    return a5 * beta + a0  # :358


def rrdb(p: Params, prefix: str, x, beta: float):
    """ResInResDenseBlock.forward, srgan_train.py:393-404."""
    a = x
    for r in (1, 2, 3):
        a = residual_dense_block(p, f"{prefix}/residual_dense_block{r}", a, beta)
    return a * beta + x  # :402


def upsample_nearest2(x):
    """F.resize_images(mode='nearest') to exactly 2x (srgan_train.py:556-566):
    out[i, j] = in[i // 2, j // 2]  (SURVEY App. B.5)."""
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def deformable_conv2d(x, offset, W, b, quantize=None):
    """chainer.functions.deformable_convolution_2d_sampler with ksize 3, stride 1, pad 1
    (call sites srgan_train.py:506-523, 572-574; semantics SURVEY App. B.6).

    offset: (N, 18, H, W); channels [0:9] are x-displacements, [9:18] y-displacements of
    tap t = ky*3 + kx.  Sampling is bilinear on the zero-padded input, zero outside.
    Written with plain index arithmetic so autograd yields d/dx, d/doffset, d/dW, d/db.
    """
    N, C, H, Wd = x.shape
    O = W.shape[0]
    dev, dt = x.device, x.dtype
    ys = torch.arange(H, device=dev, dtype=dt).view(1, 1, H, 1)
    xs = torch.arange(Wd, device=dev, dtype=dt).view(1, 1, 1, Wd)
    kx = torch.tensor([0, 1, 2] * 3, device=dev, dtype=dt).view(1, 9, 1, 1)
    ky = torch.tensor([0, 0, 0, 1, 1, 1, 2, 2, 2], device=dev, dtype=dt).view(1, 9, 1, 1)
    # position in the ORIGINAL (unpadded) frame: padded-frame coordinate minus pad (=1)
    px = xs + kx - 1.0 + offset[:, :9]
    py = ys + ky - 1.0 + offset[:, 9:]
    # Chainer clips the coordinate to one pixel outside the padded image; everything at
    # or beyond that is zero, so clipping to [-2, size+1] in this frame is equivalent.
    px = px.clamp(-2.0, Wd + 1.0)
    py = py.clamp(-2.0, H + 1.0)
    x0 = torch.floor(px)
    y0 = torch.floor(py)
    fx = px - x0
    fy = py - y0
    x0 = x0.long()
    y0 = y0.long()
    flat = x.reshape(N, C, H * Wd)

    def corner(yi, xi):
        valid = ((yi >= 0) & (yi < H) & (xi >= 0) & (xi < Wd)).to(dt)  # (N,9,H,W)
        idx = (yi.clamp(0, H - 1) * Wd + xi.clamp(0, Wd - 1)).view(N, 1, -1).expand(N, C, -1)
        v = torch.gather(flat, 2, idx).view(N, C, 9, H, Wd)
        return v * valid.unsqueeze(1)

    sampled = (corner(y0, x0) * ((1 - fy) * (1 - fx)).unsqueeze(1)
               + corner(y0, x0 + 1) * ((1 - fy) * fx).unsqueeze(1)
               + corner(y0 + 1, x0) * (fy * (1 - fx)).unsqueeze(1)
               + corner(y0 + 1, x0 + 1) * (fy * fx).unsqueeze(1))  # (N,C,9,H,W)
    if quantize is not None:  # tensor-core path: sampled operand and filter rounded to bf16
        sampled, W = quantize(sampled), quantize(W)
    y = torch.einsum("nckhw,ock->nohw", sampled, W.reshape(O, C, 9))
    return y + b.view(1, O, 1, 1)


def deformable_conv2d_fast(x, offset, W, b):
    """Same function as ``deformable_conv2d`` (forward only, no autograd use), arranged for speed: used where the
    oracle is TIMED (bench.py's CPU baseline) and for full-size 288x288 tiles, where the form above would
    materialise four (N, C, 9, H, W) gathers. Pixels are rows of a channels-last matrix (one extra all-zero row
    stands for "outside the padded image"), each tap gathers its four bilinear corners with ``index_select`` and is
    contracted by one GEMM. Same coordinate arithmetic (no [-1, 1] normalisation), so it agrees with the form
    above to rounding; equality of the two and of torchvision's deform_conv2d is asserted in
    tests/test_oracle_kat.py."""
    N, C, H, Wd = x.shape
    Oc = W.shape[0]
    dt = x.dtype
    ys = torch.arange(H, dtype=dt).view(H, 1)
    xs = torch.arange(Wd, dtype=dt).view(1, Wd)
    out = []
    for n in range(N):
        rows = torch.cat([x[n].permute(1, 2, 0).reshape(H * Wd, C), torch.zeros(1, C, dtype=dt)], 0)
        acc = torch.zeros(H * Wd, Oc, dtype=dt)
        for t in range(9):
            ky, kx = divmod(t, 3)
            px = (xs + (kx - 1.0) + offset[n, t]).clamp(-2.0, Wd + 1.0)
            py = (ys + (ky - 1.0) + offset[n, 9 + t]).clamp(-2.0, H + 1.0)
            x0, y0 = torch.floor(px), torch.floor(py)
            fx, fy = (px - x0).reshape(-1, 1), (py - y0).reshape(-1, 1)
            x0, y0 = x0.long(), y0.long()

            def corner(yi, xi):
                ok = (yi >= 0) & (yi < H) & (xi >= 0) & (xi < Wd)
                return rows.index_select(0, torch.where(ok, yi * Wd + xi, torch.full_like(yi, H * Wd)).reshape(-1))

            smp = (corner(y0, x0) * ((1 - fy) * (1 - fx)) + corner(y0, x0 + 1) * ((1 - fy) * fx)
                   + corner(y0 + 1, x0) * (fy * (1 - fx)) + corner(y0 + 1, x0 + 1) * (fy * fx))
            acc.addmm_(smp, W[:, :, ky, kx].t())
        out.append(acc.t().reshape(Oc, H, Wd))
    return torch.stack(out) + b.view(1, Oc, 1, 1)


def deformable_layer(p: Params, name: str, x):
    """L.DeformableConvolution2D = offset_conv (64->18, k3 p1, bias) + sampler."""
    off = F.conv2d(x, p[f"{name}/offset_conv/W"], p[f"{name}/offset_conv/b"], padding=1)
    return deformable_conv2d(x, off, p[f"{name}/deform_conv/W"], p[f"{name}/deform_conv/b"])


def _q(t):
    """Round to bfloat16 and back (emulates the tensor-core path's operand / storage rounding)."""
    return t.to(torch.bfloat16).to(t.dtype)


def generator_forward(p: Params, x, w1, w2, w3, num_residual_blocks=12,
                      residual_scaling=0.1, return_intermediates=False, emulate_bf16=False, fast_deform=False):
    """GeneratorModel.forward, srgan_train.py:525-576.

    ``emulate_bf16=True`` restates the SAME graph with the operand rounding of the product's
    tcgen05 path (documented in DESIGN.md): 3x3-conv inputs and weights rounded to bf16, fp32
    (here: wider) accumulation, fp32 bias / residual stream, dense-block features a1..a4 and the
    upsample-conv outputs stored as bf16; first deformable layer: bilinear samples and filter
    rounded to bf16, output stored as bf16; stem and the 64->1 output layer in fp32.
    It exists so that the CUDA kernels can be checked tightly (accumulation order is then the only
    difference) independently of how strongly a given weight set amplifies rounding noise.
    """
    if not emulate_bf16:
        return _generator_forward_exact(p, x, w1, w2, w3, num_residual_blocks, residual_scaling,
                                        return_intermediates, fast_deform=fast_deform)
    beta = residual_scaling
    q = _q

    def conv(inp_q, name):  # inp_q already bf16-representable
        return F.conv2d(inp_q, q(p[f"{name}/W"]), p[f"{name}/b"], padding=1)

    a0 = input_block(p, x, w1, w2, w3)
    a1 = _lrelu(conv(q(a0), "pre_residual_conv_layer"))          # fp32 master
    cur = a1
    for i in range(num_residual_blocks):
        rrdb_in = cur
        for r in (1, 2, 3):
            pre = f"residual_network/{i}/residual_dense_block{r}"
            feats = [q(cur)]
            for k in (1, 2, 3, 4):
                feats.append(q(_lrelu(conv(torch.cat(feats, dim=1), f"{pre}/conv_layer{k}"))))
            a5 = conv(torch.cat(feats, dim=1), f"{pre}/conv_layer5")
            cur = cur + beta * a5
            if r == 3:
                cur = rrdb_in + beta * cur
    a3 = a1 + conv(q(cur), "post_residual_conv_layer")
    u1 = upsample_nearest2(q(a3))
    c1 = q(_lrelu(conv(u1, "post_upsample_conv_layer_1")))
    c2 = q(_lrelu(conv(upsample_nearest2(c1), "post_upsample_conv_layer_2")))
    off1 = conv(c2, "final_conv_layer1/offset_conv")
    d1 = q(_lrelu(deformable_conv2d(c2, off1, p["final_conv_layer1/deform_conv/W"],
                                    p["final_conv_layer1/deform_conv/b"], quantize=q)))
    off2 = conv(d1, "final_conv_layer2/offset_conv")
    # output layer: CUDA-core dot product, fp32 filter, no operand rounding
    return deformable_conv2d(d1, off2, p["final_conv_layer2/deform_conv/W"],
                             p["final_conv_layer2/deform_conv/b"])


def trunk_forward(p: Params, a0, num_residual_blocks=12, residual_scaling=0.1, emulate_bf16=False,
                  return_a1=False):
    """a3 = a1 + post_res(RRDB^nb(a1)), a1 = lrelu(pre_res(a0)): srgan_train.py:541-551.

    ``emulate_bf16=True`` restates the SAME graph with the operand rounding of the product's tensor-core
    TRAINING trunk (deepbedmap_b200/flat.py): every 3x3-conv input and filter rounded to bf16 (dense-block
    features a1..a4 are stored as bf16), fp32 (here: wider) accumulation, bias and residual stream.
    ``_q`` is a dtype round trip, so autograd treats the rounding as straight-through -- exactly what the
    product's backward does -- and LeakyReLU derivatives are taken at the ROUNDED forward's values. This
    matters: a forward perturbation of relative size e flips the sign of a fraction ~e of the
    pre-activations, each flip changes that element's gradient by a factor 5, so gradients of a
    bf16-operand forward (e ~ 3e-3) differ from the exact graph's by ~sqrt(e) * 0.8 ~ 4e-2 in relative L2
    whatever the implementation; against this emulation only accumulation order and the bf16 rounding
    of the gradient operands remain."""
    q = _q if emulate_bf16 else (lambda t: t)
    beta = residual_scaling

    def conv(inp, name):
        return F.conv2d(inp, q(p[f"{name}/W"]), p[f"{name}/b"], padding=1)

    a1 = _lrelu(conv(q(a0), "pre_residual_conv_layer"))                      # :541-542
    cur = a1
    for i in range(num_residual_blocks):                                     # :546
        rrdb_in = cur
        for r in (1, 2, 3):
            pre = f"residual_network/{i}/residual_dense_block{r}"
            feats = [q(cur)]
            for k in (1, 2, 3, 4):                                           # :339-352
                feats.append(q(_lrelu(conv(torch.cat(feats, dim=1), f"{pre}/conv_layer{k}"))))
            cur = conv(torch.cat(feats, dim=1), f"{pre}/conv_layer5") * beta + cur   # :354-358
        cur = cur * beta + rrdb_in                                           # :402
    a3 = a1 + conv(q(cur), "post_residual_conv_layer")                       # :550-551
    return (a3, a1) if return_a1 else a3


def _generator_forward_exact(p: Params, x, w1, w2, w3, num_residual_blocks=12,
                             residual_scaling=0.1, return_intermediates=False, trunk_bf16=False, fast_deform=False):
    """GeneratorModel.forward, srgan_train.py:525-576. ``trunk_bf16``: the trunk with the operand rounding
    of the product's tensor-core training path (see trunk_forward), everything else exact."""
    inter = {}
    a0 = input_block(p, x, w1, w2, w3)                                      # :537
    if trunk_bf16:
        a3, a1 = trunk_forward(p, a0, num_residual_blocks, residual_scaling, emulate_bf16=True, return_a1=True)
        a2 = None
    else:
        a1 = _lrelu(F.conv2d(a0, p["pre_residual_conv_layer/W"],
                             p["pre_residual_conv_layer/b"], padding=1))        # :541-542
        a2 = a1
        for i in range(num_residual_blocks):                                     # :546
            a2 = rrdb(p, f"residual_network/{i}", a2, residual_scaling)
        a3 = a1 + F.conv2d(a2, p["post_residual_conv_layer/W"],
                           p["post_residual_conv_layer/b"], padding=1)          # :550-551
    # trunk_bf16 also rounds the operands of the head's four plain 3x3 convolutions (upsample convs, offset convs),
    # which the product's training path runs on the tensor cores too; the deformable sampling + contraction is fp32
    q = _q if trunk_bf16 else (lambda t: t)

    def conv(x_, name):
        return F.conv2d(q(x_), q(p[f"{name}/W"]), p[f"{name}/b"], padding=1)

    def deform(x_, name):
        fn = deformable_conv2d_fast if fast_deform else deformable_conv2d
        return fn(x_, conv(x_, f"{name}/offset_conv"), p[f"{name}/deform_conv/W"], p[f"{name}/deform_conv/b"])

    a4_1 = _lrelu(conv(upsample_nearest2(a3), "post_upsample_conv_layer_1"))    # :556-560
    a4_2 = _lrelu(conv(upsample_nearest2(a4_1), "post_upsample_conv_layer_2"))  # :562-568
    a5_1 = _lrelu(deform(a4_2, "final_conv_layer1"))                            # :572-573
    a5_2 = deform(a5_1, "final_conv_layer2")                                    # :574
    if return_intermediates:
        inter.update(a0=a0, a1=a1, a2=a2, a3=a3, a4_1=a4_1, a4_2=a4_2, a5_1=a5_1)
        return a5_2, inter
    return a5_2


# --------------------------------------------------------------------------------------
# Discriminator (srgan_train.py:591-699)
# --------------------------------------------------------------------------------------
BN_EPS = 1e-5      # :636
BN_DECAY = 0.9     # chainer default


def batch_norm(p: Params, name: str, x, train: bool, stats_out: Optional[dict]):
    """L.BatchNormalization(axis=(0,2,3), eps=1e-5).
    train: batch mean / biased variance; running stats updated with the unbiased variance.
    eval: running stats.  New running stats are returned through ``stats_out`` (functional)."""
    g = p[f"{name}/gamma"].view(1, -1, 1, 1)
    b = p[f"{name}/beta"].view(1, -1, 1, 1)
    if train:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if stats_out is not None:
            m = x.numel() // x.shape[1]
            adjust = m / max(m - 1.0, 1.0)
            stats_out[f"{name}/avg_mean"] = (BN_DECAY * p[f"{name}/avg_mean"]
                                             + (1 - BN_DECAY) * mean.detach())
            stats_out[f"{name}/avg_var"] = (BN_DECAY * p[f"{name}/avg_var"]
                                            + (1 - BN_DECAY) * var.detach() * adjust)
    else:
        mean = p[f"{name}/avg_mean"]
        var = p[f"{name}/avg_var"]
    xhat = (x - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + BN_EPS)
    return xhat * g + b


def discriminator_forward(p: Params, x, train: bool = True, stats_out: Optional[dict] = None,
                          emulate_bf16: bool = False):
    """DiscriminatorModel.forward, srgan_train.py:649-699.  Returns logits (N,1).
    ``emulate_bf16``: operand rounding of the product's tensor-core discriminator (precision="bf16"): inputs and
    filters of conv_layer1..9 rounded to bf16 (straight-through for autograd), everything else exact."""
    q = _q if emulate_bf16 else (lambda t: t)
    a = _lrelu(F.conv2d(x, p["conv_layer0/W"], p["conv_layer0/b"], stride=1, padding=1))
    for i in range(1, 10):
        _, k, s = DISC_CONVS[i]
        a = F.conv2d(q(a), q(p[f"conv_layer{i}/W"]), None, stride=s, padding=1)
        a = _lrelu(batch_norm(p, f"batch_norm{i}", a, train, stats_out))
    a = a.reshape(a.shape[0], -1)                                             # :693
    a = _lrelu(F.linear(a, p["linear_1/W"], p["linear_1/b"]))                # :694-695
    return F.linear(a, p["linear_2/W"], p["linear_2/b"])                     # :696


# --------------------------------------------------------------------------------------
# Losses and metrics (srgan_train.py:841-1009)
# --------------------------------------------------------------------------------------
def sigmoid_cross_entropy(x, t):
    """F.sigmoid_cross_entropy (normalize=True, mean over elements); formula at :980."""
    t = t.to(x.dtype)
    loss = -(x * (t - (x >= 0).to(x.dtype)) - torch.log1p(torch.exp(-torch.abs(x))))
    return loss.mean()


def calculate_discriminator_loss(real_pred, fake_pred, real_minus_fake_target,
                                 fake_minus_real_target):
    """srgan_train.py:960-1009 (RaGAN)."""
    real_avg = real_pred.mean()
    fake_avg = fake_pred.mean()
    return (sigmoid_cross_entropy(real_pred - fake_avg, real_minus_fake_target)
            + sigmoid_cross_entropy(fake_pred - real_avg, fake_minus_real_target))


def _gaussian_window(size: int, sigma: float, dtype):
    g = torch.tensor([math.exp(-((i - size // 2) ** 2) / (2.0 * sigma ** 2)) for i in range(size)],
                     dtype=dtype)
    g = g / g.sum()
    return (g[:, None] @ g[None, :])


def ssim(y_pred, y_true, window_size: int = 9, sigma: float = 1.5):
    """ssim.functions.ssim_loss of ssim-chainer@9c54f25 (pytorch-ssim lineage) as called at
    srgan_train.py:953-955: Gaussian window (sigma 1.5), VALID depth-wise convolution,
    C1 = 0.01^2, C2 = 0.03^2, mean over the whole map (KAT :944-948)."""
    if y_pred.shape != y_true.shape:
        raise ValueError("Input images must have the same dimensions.")  # :950-951
    C = y_pred.shape[1]
    win = _gaussian_window(window_size, sigma, y_pred.dtype).to(y_pred.device)
    win = win.expand(C, 1, window_size, window_size).contiguous()
    mu1 = F.conv2d(y_pred, win, groups=C)
    mu2 = F.conv2d(y_true, win, groups=C)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = F.conv2d(y_pred * y_pred, win, groups=C) - mu1_sq
    s2 = F.conv2d(y_true * y_true, win, groups=C) - mu2_sq
    s12 = F.conv2d(y_pred * y_true, win, groups=C) - mu12
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return ssim_map.mean()


def psnr(y_pred, y_true, data_range=2 ** 32):
    """srgan_train.py:906-928."""
    mse = torch.mean((y_pred - y_true) ** 2)
    return 20.0 * torch.log10(data_range / torch.sqrt(mse))


def calculate_generator_loss(y_pred, y_true, fake_labels, real_labels,
                             fake_minus_real_target, real_minus_fake_target, x_topo,
                             content_loss_weighting=1e-2, adversarial_loss_weighting=2e-2,
                             topographic_loss_weighting=2e-3, structural_loss_weighting=5.25):
    """srgan_train.py:841-902."""
    content = torch.mean(torch.abs(y_pred - y_true))                               # :871
    adversarial = calculate_discriminator_loss(real_labels, fake_labels,
                                               real_minus_fake_target,
                                               fake_minus_real_target)           # :874-879
    topo = torch.mean(torch.abs(F.avg_pool2d(y_pred, 4) - x_topo))                 # :882-884
    structural = 1 - ssim(y_pred, y_true)                                          # :887
    return (content_loss_weighting * content + adversarial_loss_weighting * adversarial
            + topographic_loss_weighting * topo + structural_loss_weighting * structural)


def binary_accuracy(y, t):
    """F.binary_accuracy: mean([y >= 0] == t) (srgan_train.py:1158)."""
    return ((y >= 0).to(torch.int64) == t.to(torch.int64)).to(torch.float64).mean()


# --------------------------------------------------------------------------------------
# Adam (Chainer variant) and the two step functions (srgan_train.py:1014-1263)
# --------------------------------------------------------------------------------------
class ChainerAdam:
    """chainer.optimizers.Adam(alpha, beta1=0.9, beta2=0.999, eps):
    m += (1-b1)(g-m); v += (1-b2)(g^2-v); p -= alpha*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps)."""

    def __init__(self, alpha=1.6e-4, beta1=0.9, beta2=0.999, eps=1e-8):
        self.alpha, self.beta1, self.beta2, self.eps = alpha, beta1, beta2, eps
        self.t = 0
        self.m: Dict[str, torch.Tensor] = {}
        self.v: Dict[str, torch.Tensor] = {}

    def update(self, params: Params, grads: Dict[str, torch.Tensor]):
        self.t += 1
        fix1 = 1.0 - self.beta1 ** self.t
        fix2 = 1.0 - self.beta2 ** self.t
        lr = self.alpha * math.sqrt(fix2) / fix1
        with torch.no_grad():
            for k, g in grads.items():
                if g is None:
                    continue
                if k not in self.m:
                    self.m[k] = torch.zeros_like(params[k])
                    self.v[k] = torch.zeros_like(params[k])
                self.m[k] += (1 - self.beta1) * (g - self.m[k])
                self.v[k] += (1 - self.beta2) * (g * g - self.v[k])
                params[k] -= lr * self.m[k] / (torch.sqrt(self.v[k]) + self.eps)


def _trainable(p: Params):
    return {k: v for k, v in p.items()
            if v.is_floating_point() and not (k.endswith("avg_mean") or k.endswith("avg_var"))}


def train_eval_discriminator(arrays: Dict[str, torch.Tensor], g_params: Params, d_params: Params,
                             d_optimizer: Optional[ChainerAdam] = None, train: bool = True,
                             num_residual_blocks=12, residual_scaling=0.1, return_grads=False):
    """srgan_train.py:1084-1166.  Mutates d_params (weights + BN running stats) in place."""
    if train:
        assert d_optimizer is not None or return_grads                        # :1127
    with torch.no_grad():                                                      # :1131
        fake = generator_forward(g_params, arrays["X"], arrays["W1"], arrays["W2"], arrays["W3"],
                                 num_residual_blocks, residual_scaling)
    real = arrays["Y"]
    n = real.shape[0]
    tr = _trainable(d_params)
    for v in tr.values():
        v.requires_grad_(train)
    stats: dict = {}
    # NOTE: two separate BN-statistics passes, real first then fake (:1145-1146); the
    # running stats are updated twice in that order.
    real_pred = discriminator_forward(d_params, real, train=train, stats_out=stats)
    if train:
        with torch.no_grad():
            for k, v in stats.items():
                d_params[k] = v
    stats2: dict = {}
    fake_pred = discriminator_forward(d_params, fake, train=train, stats_out=stats2)
    if train:
        with torch.no_grad():
            for k, v in stats2.items():
                d_params[k] = v
            for i in range(1, 10):
                d_params[f"batch_norm{i}/N"] = d_params[f"batch_norm{i}/N"] + 2
    ones = torch.ones(n, 1, dtype=torch.int64)
    zeros = torch.zeros(n, 1, dtype=torch.int64)
    d_loss = calculate_discriminator_loss(real_pred, fake_pred, ones, zeros)  # :1149-1154
    pred = torch.cat([real_pred.detach(), fake_pred.detach()])
    d_accu = binary_accuracy(pred, torch.cat([ones, zeros]))                   # :1156-1158
    grads = None
    if train:
        gl = torch.autograd.grad(d_loss, list(tr.values()), allow_unused=True)  # :1162-1163
        grads = {k: g for k, g in zip(tr.keys(), gl)}
        for v in tr.values():
            v.requires_grad_(False)
        if d_optimizer is not None:
            d_optimizer.update(d_params, grads)                                 # :1164
    out = (float(d_loss.detach()), float(d_accu))
    return out + (grads,) if return_grads else out


def train_eval_generator(arrays: Dict[str, torch.Tensor], g_params: Params, d_params: Params,
                         g_optimizer: Optional[ChainerAdam] = None, train: bool = True,
                         num_residual_blocks=12, residual_scaling=0.1, return_grads=False):
    """srgan_train.py:1170-1263.  Mutates g_params in place."""
    if train:
        assert g_optimizer is not None or return_grads                        # :1218
    tr = _trainable(g_params)
    for v in tr.values():
        v.requires_grad_(train)
    fake = generator_forward(g_params, arrays["X"], arrays["W1"], arrays["W2"], arrays["W3"],
                             num_residual_blocks, residual_scaling)             # :1222-1227
    with torch.no_grad():                                                       # :1228-1229
        # eval-mode BN and `.array` => detached: no gradient flows through D
        fake_labels = discriminator_forward(d_params, fake.detach(), train=False)
    real = arrays["Y"]
    n = real.shape[0]
    real_labels = torch.ones(n, 1, dtype=fake.dtype)                            # :1233
    ones = torch.ones(n, 1, dtype=torch.int64)
    zeros = torch.zeros(n, 1, dtype=torch.int64)
    g_loss = calculate_generator_loss(fake, real, fake_labels, real_labels,
                                      fake_minus_real_target=ones,
                                      real_minus_fake_target=zeros,
                                      x_topo=arrays["X"][:, :, 1:-1, 1:-1])    # :1238-1249
    g_psnr = psnr(fake.detach(), real)                                          # :1250
    g_ssim = ssim(fake.detach(), real)                                          # :1251
    grads = None
    if train:
        gl = torch.autograd.grad(g_loss, list(tr.values()), allow_unused=True)  # :1255-1256
        grads = {k: g for k, g in zip(tr.keys(), gl)}
        for v in tr.values():
            v.requires_grad_(False)
        if g_optimizer is not None:
            g_optimizer.update(g_params, grads)                                 # :1257
    out = (float(g_loss.detach()), float(g_psnr), float(g_ssim))
    return out + (grads,) if return_grads else out


# --------------------------------------------------------------------------------------
# Continent tiler (deepbedmap.py:681-740)
# --------------------------------------------------------------------------------------
def tile_plan(final_shape=(18000, 22000), ary_shape=(1000, 1000), stride=(1000, 1000),
              xtrapad=(18, 18)):
    """Yields (y0, y1, x0, x1, y_slice, x_slice) exactly as deepbedmap.py:700-732.
    Coordinates are lowres (BEDMAP2) pixels; slices index the 4x output grid."""
    plan = []
    for sy in range(0, final_shape[0], stride[0]):
        for sx in range(0, final_shape[1], stride[1]):
            y0 = max(0, (sy // 4) - xtrapad[0] - 1)
            y1 = min(final_shape[0] // 4, ((sy + ary_shape[0]) // 4) + xtrapad[0] + 1)
            x0 = max(0, (sx // 4) - xtrapad[1] - 1)
            x1 = min(final_shape[1] // 4, ((sx + ary_shape[1]) // 4) + xtrapad[1] + 1)
            ys = slice((y0 + xtrapad[0] + 1) * 4, (y1 - xtrapad[0] - 1) * 4)
            xs = slice((x0 + xtrapad[1] + 1) * 4, (x1 - xtrapad[1] - 1) * 4)
            plan.append((y0, y1, x0, x1, ys, xs))
    return plan


def predict_continent(forward: Callable, X, W1, W2, W3, final_shape=(18000, 22000),
                      ary_shape=(1000, 1000), stride=(1000, 1000), xtrapad=(18, 18)):
    """deepbedmap.py:696-740: NaN-filled canvas, batch-1 tiles, crop `xtrapad*4` px borders.
    ``forward(x, w1, w2, w3) -> ndarray (1,1,H,W)``; inputs are NumPy (1,C,h,w) arrays."""
    Y_hat = np.full((1, final_shape[0], final_shape[1]), np.nan, dtype=np.float32)
    for (y0, y1, x0, x1, ys, xs) in tile_plan(final_shape, ary_shape, stride, xtrapad):
        xc = np.asarray(X[:, :, y0:y1, x0:x1], dtype=np.float32)
        w1c = np.asarray(W1[:, :, y0 * 10:y1 * 10, x0 * 10:x1 * 10], dtype=np.float32)
        w2c = np.asarray(W2[:, :, y0 * 2:y1 * 2, x0 * 2:x1 * 2], dtype=np.float32)
        w3c = np.asarray(W3[:, :, y0:y1, x0:x1], dtype=np.float32)
        y_pred = np.asarray(forward(xc, w1c, w2c, w3c))[0]
        py, px = xtrapad[0] * 4, xtrapad[1] * 4
        Y_hat[:, ys, xs] = y_pred[:, py:-py, px:-px]
    return Y_hat


def synthetic_continent(grid=(4502, 5502), seed: int = 42):
    """SURVEY 8(d) config 3 inputs on the host, physical regime: X (1,1,H,W) BEDMAP2-like metres,
    W1 (1,1,10H,10W) REMA-like metres, W2 (1,2,2H,2W) velocities (signed: the tiler's clip matters),
    W3 (1,1,H,W). NumPy RandomState, so the same values on every machine."""
    H, W = grid
    rng = np.random.RandomState(seed)
    X = np.clip(rng.normal(-500.0, 800.0, (1, 1, H, W)), -5000.0, 4500.0)
    W1 = rng.uniform(0.0, 4000.0, (1, 1, 10 * H, 10 * W))
    W2 = rng.normal(0.0, 200.0, (1, 2, 2 * H, 2 * W))
    W3 = rng.uniform(-50.0, 1000.0, (1, 1, H, W))
    return tuple(a.astype(np.float32) for a in (X, W1, W2, W3))


def continent_tile_inputs(X, W1, W2, W3, tile):
    """The four crops the reference feeds the generator for one tile of ``tile_plan`` (deepbedmap.py:715-722),
    W1..W3 clipped at zero as the reference does for the whole arrays beforehand (deepbedmap.py:663-665)."""
    y0, y1, x0, x1 = tile[:4]
    return (np.ascontiguousarray(X[:, :, y0:y1, x0:x1], dtype=np.float32),
            np.clip(W1[:, :, y0 * 10:y1 * 10, x0 * 10:x1 * 10], 0, None).astype(np.float32),
            np.clip(W2[:, :, y0 * 2:y1 * 2, x0 * 2:x1 * 2], 0, None).astype(np.float32),
            np.clip(W3[:, :, y0:y1, x0:x1], 0, None).astype(np.float32))


# --------------------------------------------------------------------------------------
# Convenience wrappers used by tests / bench
# --------------------------------------------------------------------------------------
def generator_forward_numpy(params_np, x, w1, w2, w3, num_residual_blocks=12,
                            residual_scaling=0.1, dtype=torch.float64, emulate_bf16=False,
                            fast_deform=False) -> np.ndarray:
    p = to_torch(params_np, dtype)
    with torch.no_grad():
        y = generator_forward(p, *(torch.as_tensor(np.asarray(a)).to(dtype) for a in (x, w1, w2, w3)),
                              num_residual_blocks=num_residual_blocks,
                              residual_scaling=residual_scaling, emulate_bf16=emulate_bf16, fast_deform=fast_deform)
    return y.numpy()


def synthetic_inputs(n: int, h: int = 11, w: int = 11, seed: int = 42, regime: str = "unit"):
    """'unit' = RandomState(seed).rand in the order X, W1, W2, W3
    (doctest regime, srgan_train.py:439-442); 'physical' = metre-scale values."""
    rng = np.random.RandomState(seed)
    if regime == "unit":
        X = rng.rand(n, 1, h, w)
        W1 = rng.rand(n, 1, 10 * h, 10 * w)
        W2 = rng.rand(n, 2, 2 * h, 2 * w)
        W3 = rng.rand(n, 1, h, w)
    elif regime == "physical":
        X = np.clip(rng.normal(-500.0, 800.0, (n, 1, h, w)), -5000.0, 4500.0)
        W1 = rng.uniform(0.0, 4000.0, (n, 1, 10 * h, 10 * w))
        W2 = rng.normal(0.0, 200.0, (n, 2, 2 * h, 2 * w))
        W3 = rng.uniform(0.0, 1000.0, (n, 1, h, w))
    else:
        raise ValueError(regime)
    return tuple(a.astype(np.float32) for a in (X, W1, W2, W3))
