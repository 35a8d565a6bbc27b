"""Small cases of every kernel that synchronises through mbarriers / global flags / proxy fences, for
compute-sanitizer (scripts/sanitize.sh). Each case runs the product path through the C ABI on shapes small enough for
the sanitizer's 10-100x slowdown and checks the result against the exact fp32 CUDA path, so a sanitizer-clean run
is also a correct one.

usage: python scripts/sanitize_cases.py [trunk|split|local|train|chain|chain64|all]
  trunk : dbm_stem_w1_s2d, dbm_conv3x3_umma(_valid), dbm_trunk_umma (persistent, paired plan, 2 x 21x38 -> 3 units per
          pass and image), dbm_deform_conv_umma, dbm_deform_conv_out1
  split : precision="bf16x3": dbm_nchw_to_slab8_split, dbm_trunk_umma_split (trunk + both upsample convs), dbm_slab8f_to_nchw
  local : dbm_stem_fwd_flat, dbm_trunk_local_fwd (image-resident trunk, 3 images of 11x11)
  train : one D-step + G-step at batch 3, 1 RRDB: dbm_trunk_local_fwd/bwd, dbm_flat_conv3x3_seq, dbm_flat_wgrad(+reduce),
          BatchNorm, losses, Adam
  chain : the same step with the image-resident kernels off: dbm_flat_conv3x3_chain (flag-synchronised layer chain)
  chain64: the step at inter_channels = 64 (flat chain, data gradients of conv4 / conv5 as N-slices)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepbedmap_b200 import GeneratorModel, flat  # noqa: E402
from deepbedmap_b200 import train as T  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def inputs(n, h, w, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g, device="cuda")
    return r(n, 1, h, w), r(n, 1, 10 * h, 10 * w), r(n, 2, 2 * h, 2 * w), r(n, 1, h, w)


def forward_case(n, h, w, local):
    m16 = GeneratorModel(num_residual_blocks=1, precision="bf16", seed=0, init_scale=0.7)
    m32 = GeneratorModel(num_residual_blocks=1, precision="fp32", seed=0, init_scale=0.7)
    m16.local_trunk = local
    ins = inputs(n, h, w)
    y16, y32 = m16.forward(*ins).array, m32.forward(*ins).array
    torch.cuda.synchronize()
    e = rel(y16, y32)
    print(f"forward {n}x{h}x{w} local={local}: bf16 vs fp32 rel_l2 {e:.3e}")
    assert e < 2e-2


def split_case(n, h, w):
    m3 = GeneratorModel(num_residual_blocks=1, precision="bf16x3", seed=0, init_scale=0.7)
    m32 = GeneratorModel(num_residual_blocks=1, precision="fp32", seed=0, init_scale=0.7)
    ins = inputs(n, h, w)
    y3, y32 = m3.forward(*ins).array, m32.forward(*ins).array
    torch.cuda.synchronize()
    e = rel(y3, y32)
    print(f"forward {n}x{h}x{w} bf16x3 vs fp32 rel_l2 {e:.3e}")
    assert e < 1e-4


def train_case(local, inter_channels=32):
    saved = flat.local_trunk_fits
    if not local:
        flat.local_trunk_fits = lambda H, W: False
    try:
        g, g_opt, d, d_opt = T.compile_srgan_model(num_residual_blocks=1, inter_channels=inter_channels)
        gen = torch.Generator(device="cuda").manual_seed(1)
        r = lambda *s: torch.rand(*s, generator=gen, device="cuda")
        arrays = {"X": r(3, 1, 11, 11), "W1": r(3, 1, 110, 110), "W2": r(3, 2, 22, 22), "W3": r(3, 1, 11, 11),
                  "Y": r(3, 1, 36, 36)}
        w0 = g.flat.clone()
        dl, da = T.train_eval_discriminator(arrays, g, d, d_opt, share_generator_forward=True)
        gl, gp, gs = T.train_eval_generator(arrays, g, d, g_opt)
        torch.cuda.synchronize()
        print(f"train step local={local} inter_channels={inter_channels}: d_loss {dl:.5f} g_loss {gl:.5f} psnr {gp:.3f} ssim {gs:.5f}")
        assert all(np.isfinite(v) for v in (dl, da, gl, gp, gs)) and not torch.equal(w0, g.flat)
    finally:
        flat.local_trunk_fits = saved


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("trunk", "all"):
        forward_case(2, 21, 38, local=False)
    if which in ("local", "all"):
        forward_case(3, 11, 11, local=True)
    if which in ("train", "all"):
        train_case(local=True)
    if which in ("chain", "all"):
        train_case(local=False)
    if which in ("chain64", "all"):
        train_case(local=False, inter_channels=64)
    if which in ("split", "all"):
        split_case(2, 21, 38)
    print("sanitize cases OK")


if __name__ == "__main__":
    main()
