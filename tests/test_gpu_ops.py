"""GPU parity of every C-ABI kernel against plain torch / the CPU oracle on seeded inputs.
fp32 kernels: relative L2 <= 2e-6 (accumulation-order noise only).
tcgen05 (bf16 operand) kernel: compared against an fp64 convolution of the SAME bf16-rounded
operands, so only fp32 accumulation order (+ bf16 output rounding where the output is bf16)
separates the two: relative L2 <= 1e-5 (fp32 out) / 4e-3 (bf16 out)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a = a.double().cpu()
    b = b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def ops():
    from deepbedmap_b200 import ops as o
    return o


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


CONV_CASES = [  # n, cin, h, w, cout, k, s, p
    (3, 5, 11, 13, 7, 3, 1, 1), (2, 1, 11, 11, 32, 3, 1, 0), (2, 1, 110, 110, 32, 30, 10, 0),
    (2, 2, 22, 26, 32, 6, 2, 0), (4, 64, 36, 36, 64, 4, 2, 1), (5, 128, 9, 9, 128, 4, 2, 1),
    (2, 192, 9, 9, 64, 3, 1, 1), (3, 512, 2, 2, 512, 4, 2, 1), (2, 70, 17, 9, 65, 3, 1, 1),
]


@pytest.mark.parametrize("n,cin,h,w,cout,k,s,p", CONV_CASES)
def test_conv2d_fwd_bwd(ops, n, cin, h, w, cout, k, s, p):
    x = rnd(n, cin + 3, h, w, seed=1)           # read a channel slice [2, 2+cin)
    wt = rnd(cout, cin, k, k, seed=2, scale=0.2)
    b = rnd(cout, seed=3)
    ho, wo = ops.conv_out_hw(h, w, k, s, p)
    y = torch.full((n, cout + 4, ho, wo), 7.0, device="cuda")
    ops.conv2d_fwd(x, 2, cin, wt, b, y, 1, k, s, p, act=True)
    xs = x[:, 2:2 + cin].double().requires_grad_(True)
    wd = wt.double().requires_grad_(True)
    ref = F.leaky_relu(F.conv2d(xs, wd, b.double(), stride=s, padding=p), 0.2)
    assert rel_l2(y[:, 1:1 + cout], ref) < 2e-6
    assert torch.all(y[:, 0] == 7.0) and torch.all(y[:, 1 + cout:] == 7.0)  # untouched channels
    # gradients of sum(conv * dy) (no activation)
    dy = rnd(n, cout, ho, wo, seed=4)
    lin = F.conv2d(xs, wd, None, stride=s, padding=p)
    gx, gw = torch.autograd.grad((lin * dy.double()).sum(), [xs, wd])
    dx = torch.full((n, cin + 3, h, w), 1.0, device="cuda")
    ops.conv2d_bwd_data(dy, 0, wt, dx, 2, cin, k, s, p, accumulate=True)
    assert rel_l2(dx[:, 2:2 + cin] - 1.0, gx) < 5e-6
    assert torch.all(dx[:, :2] == 1.0)
    dw = torch.zeros_like(wt)
    db = torch.zeros(cout, device="cuda")
    ops.conv2d_bwd_weight(x, 2, cin, dy, 0, dw, k, s, p, db=db)
    assert rel_l2(dw, gw) < 5e-6
    assert rel_l2(db, dy.double().sum(dim=(0, 2, 3))) < 5e-6


def test_gemm_variants(ops):
    a = rnd(37, 91, seed=1)
    b = rnd(91, 53, seed=2)
    bias = rnd(53, seed=3)
    c = ops.empty(37, 53)
    ops.gemm(a, 91, 1, 0, b, 53, 1, 0, c, 53, 1, 0, bias, 37, 53, 91)
    assert rel_l2(c, a.double() @ b.double() + bias.double()) < 2e-6
    # A^T (m-contiguous), B^T (k-contiguous), C^T, batched + atomic reduction
    at = a.t().contiguous()
    bt = b.t().contiguous()
    ct = ops.empty(53, 37)
    ops.gemm(at, 1, 37, 0, bt, 1, 91, 0, ct, 1, 37, 0, None, 37, 53, 91)
    assert rel_l2(ct.t(), a.double() @ b.double()) < 2e-6
    ab = rnd(4, 37, 91, seed=5)
    acc = ops.zeros(37, 53)
    ops.gemm(ab, 91, 1, 37 * 91, b, 53, 1, 0, acc, 53, 1, 0, None, 37, 53, 91, batch=4, accumulate=2)
    assert rel_l2(acc, (ab.double() @ b.double()).sum(0)) < 2e-6


@pytest.mark.parametrize("n,hw,o,k", [(3, 1296, 64, 576), (2, 100, 64, 576), (5, 81, 40, 72)])
def test_gemm_bf16_deformable_contractions(ops, n, hw, o, k):
    """dbm_gemm_bf16 on the three strided contractions of the deformable layer (forward, weight gradient summed
    over the batch, cols gradient): fp32 memory, operands rounded to bf16, fp32 accumulation."""
    g = torch.Generator().manual_seed(3)
    cols = torch.randn(n, k, hw, generator=g).cuda()
    w = (torch.randn(o, k, generator=g) * 0.1).cuda()
    b = torch.randn(o, generator=g).cuda()
    dy = torch.randn(n, o, hw, generator=g).cuda()
    q = lambda t: t.to(torch.bfloat16).double()
    # y[n, o, p] = lrelu(sum_k cols[n, k, p] w[o, k] + b[o])
    y = torch.empty(n, o, hw, device="cuda")
    ops.gemm(cols, 1, hw, k * hw, w, 1, k, 0, y, 1, hw, o * hw, b, hw, o, k, batch=n, act=True, tc=True)
    ref = torch.einsum("nkp,ok->nop", q(cols), q(w)) + b.double().view(1, o, 1)
    ref = torch.where(ref >= 0, ref, 0.2 * ref)
    assert rel_l2(y, ref) < 1e-5
    # dw[o, kk] += sum_{n, p} dy[n, o, p] cols[n, kk, p]
    dw = torch.ones(o, k, device="cuda")
    ops.gemm(dy, hw, 1, o * hw, cols, 1, hw, k * hw, dw, k, 1, 0, None, o, k, hw, batch=n, accumulate=2, tc=True)
    ref = 1.0 + torch.einsum("nop,nkp->ok", q(dy), q(cols))
    assert rel_l2(dw, ref) < 1e-5
    # dcols[n, kk, p] = sum_o w[o, kk] dy[n, o, p]
    dcols = torch.empty(n, k, hw, device="cuda")
    ops.gemm(dy, 1, hw, o * hw, w, k, 1, 0, dcols, 1, hw, k * hw, None, hw, k, o, batch=n, tc=True)
    ref = torch.einsum("ok,nop->nkp", q(w), q(dy))
    assert rel_l2(dcols, ref) < 1e-5


def test_elementwise_and_layouts(ops):
    x = rnd(3, 16, 7, 5, seed=1)
    y = rnd(3, 24, 7, 5, seed=2)
    out = torch.zeros(3, 40, 7, 5, device="cuda")
    ops.axpby(x, 4, y, 8, out, 16, 8, 0.1, 1.0)
    assert rel_l2(out[:, 16:24], 0.1 * x[:, 4:12] + y[:, 8:16]) < 1e-7
    up = ops.upsample2_fwd(x)
    assert torch.equal(up, x.repeat_interleave(2, 2).repeat_interleave(2, 3))
    dn = ops.upsample2_bwd(up)
    assert rel_l2(dn, 4 * x) < 1e-7
    yy = F.leaky_relu(x, 0.2)
    dx = torch.zeros_like(x)
    ops.lrelu_bwd(y[:, :16].contiguous(), 0, yy, 0, dx, 0, 16)
    assert rel_l2(dx, y[:, :16] * torch.where(x >= 0, 1.0, 0.2)) < 1e-7
    # slab layouts round trip
    s8 = ops.empty(3, 3, 7, 5, 8, dtype=torch.bfloat16)
    ops.nchw_to_slab8(x, s8, dst_cs0=1)
    back = ops.slab8_to_nchw(s8, 16, src_cs0=1)
    assert torch.equal(back, x.bfloat16().float())
    assert torch.equal(s8[:, 1:].float(), x.bfloat16().float().view(3, 2, 8, 7, 5).permute(0, 1, 3, 4, 2))
    s4 = ops.nchw_to_slab4(x)
    assert torch.equal(s4, x.view(3, 4, 4, 7, 5).permute(0, 1, 3, 4, 2))
    assert torch.equal(ops.slab4_to_nchw(s4, 14), x[:, :14])


@pytest.mark.parametrize("n,h,w", [(2, 11, 11), (1, 21, 37), (3, 10, 19)])
def test_fused_stem_kernel(ops, n, h, w):
    """dbm_stem_fwd_slab8 vs the four valid strided convs + concat (srgan_train.py:256-266)."""
    x = rnd(n, 1, h, w, seed=1)
    w1 = rnd(n, 1, 10 * h, 10 * w, seed=2)
    w2 = rnd(n, 2, 2 * h, 2 * w, seed=3)
    w3 = rnd(n, 1, h, w, seed=4)
    fx, f1, f2, f3 = (rnd(32, 1, 3, 3, seed=5), rnd(32, 1, 30, 30, seed=6, scale=0.05),
                      rnd(32, 2, 6, 6, seed=7, scale=0.2), rnd(32, 1, 3, 3, seed=8))
    bias = rnd(128, seed=9)
    wt1 = ops.empty(900, 32)
    ops.call("dbm_transpose_f32", f1.data_ptr(), wt1.data_ptr(), 32, 900, ops.stream())
    assert torch.equal(wt1, f1.view(32, 900).t())
    wts = torch.cat([f.reshape(32, -1).t() for f in (fx, f2, f3)]).contiguous()
    out = ops.empty(n, 16, h - 2, w - 2, 8, dtype=torch.bfloat16)
    ops.call("dbm_stem_fwd_slab8", x.data_ptr(), w1.data_ptr(), w2.data_ptr(), w3.data_ptr(), wt1.data_ptr(),
             wts.data_ptr(), bias.data_ptr(), out.data_ptr(), 16, 0, n, h, w, ops.stream())
    torch.cuda.synchronize()
    d = torch.double
    ref = torch.cat([F.conv2d(x.to(d), fx.to(d), bias[:32].to(d)), F.conv2d(w1.to(d), f1.to(d), bias[32:64].to(d), stride=10),
                     F.conv2d(w2.to(d), f2.to(d), bias[64:96].to(d), stride=2), F.conv2d(w3.to(d), f3.to(d), bias[96:].to(d))],
                    dim=1)
    got = ops.slab8_to_nchw(out, 128)
    assert got.shape == ref.shape
    assert rel_l2(got, ref) < 4e-3
    assert rel_l2(got, ref.float().bfloat16().float()) < 2e-4   # only the final bf16 rounding separates them


UMMA_CASES = [  # n, cin(read), in_cs_total, h, w, cout_real, cout_pad
    (2, 64, 8, 20, 23, 32, 32), (1, 192, 24, 37, 41, 64, 64), (3, 128, 16, 9, 9, 64, 64),
    (2, 96, 24, 16, 16, 32, 32), (1, 64, 8, 33, 18, 18, 32), (5, 160, 24, 9, 9, 32, 32),
]


@pytest.mark.parametrize("n,cin,cs,h,w,cout,coutp", UMMA_CASES)
def test_conv3x3_umma_plain(ops, n, cin, cs, h, w, cout, coutp):
    xfull = rnd(n, cs * 8, h, w, seed=1).bfloat16().float()
    wt = rnd(cout, cin, 3, 3, seed=2, scale=0.1).bfloat16().float()
    b = rnd(cout, seed=3)
    bp = torch.zeros(coutp, device="cuda")
    bp[:cout] = b
    s8 = ops.empty(n, cs, h, w, 8, dtype=torch.bfloat16)
    ops.nchw_to_slab8(xfull, s8)
    packed = ops.pack_conv3x3(wt, coutp)
    out32 = ops.empty(n, coutp // 4, h, w, 4)
    ops.conv3x3_umma(s8, cin, packed, bp, coutp, out_f32=out32)
    torch.cuda.synchronize()
    got = ops.slab4_to_nchw(out32, cout)
    ref = F.conv2d(xfull[:, :cin].double(), wt.double(), b.double(), padding=1)
    assert rel_l2(got, ref) < 1e-5
    if coutp > cout:  # padded output channels are exactly zero (zero weights, zero bias)
        assert torch.all(ops.slab4_to_nchw(out32, coutp)[:, cout:] == 0)


def test_conv3x3_umma_fused_epilogues(ops):
    n, h, w = 2, 19, 21
    cat = rnd(n, 192, h, w, seed=1).bfloat16().float()
    w1 = rnd(32, 96, 3, 3, seed=2, scale=0.1).bfloat16().float()
    b1 = rnd(32, seed=3)
    s8 = ops.empty(n, 24, h, w, 8, dtype=torch.bfloat16)
    ops.nchw_to_slab8(cat, s8)
    # (1) bias + LeakyReLU, bf16 result written into the concat slot [96, 128) of the SAME buffer
    ops.conv3x3_umma(s8, 96, ops.pack_conv3x3(w1, 32), b1, 32, act=True, out=s8, out_cs0=12)
    torch.cuda.synchronize()
    got = ops.slab8_to_nchw(s8, 192)
    ref = F.leaky_relu(F.conv2d(cat[:, :96].double(), w1.double(), b1.double(), padding=1), 0.2)
    assert rel_l2(got[:, 96:128], ref) < 4e-3
    assert torch.equal(got[:, :96], cat[:, :96]) and torch.equal(got[:, 128:], cat[:, 128:])
    # (2) conv5: residual scaling with two fp32 residual streams, fp32 + bf16 outputs
    w5 = rnd(64, 192, 3, 3, seed=4, scale=0.05).bfloat16().float()
    b5 = rnd(64, seed=5)
    r1 = rnd(n, 64, h, w, seed=6)
    r2 = rnd(n, 64, h, w, seed=7)
    nxt = ops.zeros(n, 24, h, w, 8, dtype=torch.bfloat16)
    o32 = ops.empty(n, 16, h, w, 4)
    cur = ops.empty(n, 24, h, w, 8, dtype=torch.bfloat16)
    ops.nchw_to_slab8(cat, cur)
    ops.conv3x3_umma(cur, 192, ops.pack_conv3x3(w5, 64), b5, 64, beta=0.1, out=nxt, out_f32=o32,
                     res1=ops.nchw_to_slab4(r1), res2=ops.nchw_to_slab4(r2))
    torch.cuda.synchronize()
    conv = F.conv2d(cat.double(), w5.double(), b5.double(), padding=1)
    ref = r2.double() + 0.1 * (r1.double() + 0.1 * conv)
    assert rel_l2(ops.slab4_to_nchw(o32, 64), ref) < 1e-5
    assert rel_l2(ops.slab8_to_nchw(nxt, 64), ref) < 4e-3
    # (3) nearest x2 replication fused into the store
    up = ops.empty(n, 8, 2 * h, 2 * w, 8, dtype=torch.bfloat16)
    w6 = rnd(64, 64, 3, 3, seed=8, scale=0.1).bfloat16().float()
    ops.conv3x3_umma(cur, 64, ops.pack_conv3x3(w6, 64), b5, 64, act=True, up2=True, out=up)
    torch.cuda.synchronize()
    ref = F.leaky_relu(F.conv2d(cat[:, :64].double(), w6.double(), b5.double(), padding=1), 0.2)
    ref = ref.repeat_interleave(2, 2).repeat_interleave(2, 3)
    assert rel_l2(ops.slab8_to_nchw(up, 64), ref) < 4e-3


@pytest.mark.parametrize("n,h,w", [(1, 8, 16), (2, 37, 45), (3, 36, 36)])
def test_deform_conv_tensor_core_path(ops, n, h, w):
    """dbm_deform_conv_umma (64->64) and dbm_deform_conv_out1 (64->1) vs the oracle's deformable
    convolution on the same bf16-representable input."""
    from oracle import deepbedmap_oracle as O
    x = rnd(n, 64, h, w, seed=1).bfloat16().float()
    off = rnd(n, 18, h, w, seed=2, scale=1.5)
    off[0, :, 0, 0] = 50.0
    off[0, :9, 1, 1] = -30.0
    offp = torch.zeros(n, 32, h, w, device="cuda")
    offp[:, :18] = off
    x8 = ops.empty(n, 8, h, w, 8, dtype=torch.bfloat16)
    ops.nchw_to_slab8(x, x8)
    off4 = ops.nchw_to_slab4(offp)
    wt = rnd(64, 64, 3, 3, seed=3, scale=0.1)
    b = rnd(64, seed=4)
    out = ops.empty(n, 8, h, w, 8, dtype=torch.bfloat16)
    ops.deform_conv_umma(x8, off4, ops.pack_conv3x3(wt, 64, ck=64), b, out, act=True)
    torch.cuda.synchronize()
    ref = F.leaky_relu(O.deformable_conv2d(x.double().cpu(), off.double().cpu(), wt.double().cpu(), b.double().cpu(),
                                           quantize=O._q), 0.2)
    assert rel_l2(ops.slab8_to_nchw(out, 64), ref) < 4e-3
    w1 = rnd(1, 64, 3, 3, seed=5, scale=0.1)
    b1 = rnd(1, seed=6)
    y = ops.deform_conv_out1(x8, off4, w1, b1)
    ref1 = O.deformable_conv2d(x.double().cpu(), off.double().cpu(), w1.double().cpu(), b1.double().cpu())
    assert rel_l2(y, ref1) < 1e-5


@pytest.mark.parametrize("n,h,w", [(1, 8, 16), (2, 37, 45), (3, 36, 36)])
def test_deform_sample_from_slab8_equals_fp32_sampler(ops, n, h, w):
    """dbm_deform_sample_slab8_f32 (16-byte corner gathers from the bf16 slab8 input of the fused forward) against
    dbm_deform_sample_f32 on the same bf16-representable values: the same bilinear samples, incl. out-of-image corners
    (large offsets) -- up to fp32 association."""
    x = rnd(n, 64, h, w, seed=1).bfloat16().float()
    off = rnd(n, 18, h, w, seed=2, scale=1.5)
    off[0, :, 0, 0] = 50.0
    off[0, :9, 1, 1] = -30.0
    x8 = ops.empty(n, 8, h, w, 8, dtype=torch.bfloat16)
    ops.nchw_to_slab8(x, x8)
    got = ops.deform_sample_slab8(x8, off)
    want = ops.deform_sample(x, off)
    assert tuple(got.shape) == tuple(want.shape) == (n, 576, h * w)
    assert rel_l2(got, want) < 1e-6 and float((got - want).abs().max()) < 1e-5


def test_deform_conv_fwd_bwd(ops):
    from oracle import deepbedmap_oracle as O
    n, c, h, w, o = 2, 6, 9, 11, 5
    x = rnd(n, c, h, w, seed=1)
    off = rnd(n, 18, h, w, seed=2, scale=1.5)
    off[0, :, 0, 0] = 40.0   # far outside: clamped, zero value and zero offset-gradient
    wt = rnd(o, c, 3, 3, seed=3, scale=0.3)
    b = rnd(o, seed=4)
    y, cols = ops.deform_conv_fwd(x, off, wt, b)
    xs, os_, ws, bs = (t.double().cpu().requires_grad_(True) for t in (x, off, wt, b))
    ref = O.deformable_conv2d(xs, os_, ws, bs)
    assert rel_l2(y, ref) < 2e-6
    dy = rnd(n, o, h, w, seed=5)
    gx, goff, gw, gb = torch.autograd.grad((ref * dy.double().cpu()).sum(), [xs, os_, ws, bs])
    dw, db, dx = torch.zeros_like(wt), torch.zeros_like(b), torch.zeros_like(x)
    doff = ops.deform_conv_bwd(x, off, wt, cols, dy, dw, db, dx)
    assert rel_l2(dw, gw) < 5e-6 and rel_l2(db, gb) < 5e-6
    assert rel_l2(dx, gx) < 5e-6
    assert rel_l2(doff, goff) < 5e-5


@pytest.mark.parametrize("n,c,h,w", [(2, 6, 9, 11), (3, 64, 36, 36), (1, 64, 20, 52)])
def test_deform1_tap_projection_fwd_bwd(ops, n, c, h, w):
    """Single-output deformable layer by tap projection (final_conv_layer2) vs the oracle and its autograd."""
    from oracle import deepbedmap_oracle as O
    x = rnd(n, c, h, w, seed=1)
    off = rnd(n, 18, h, w, seed=2, scale=1.5)
    off[0, :, 0, 0] = 40.0   # far outside: clamped, zero value and zero offset-gradient
    off[0, :9, 1, 1] = -3.25  # samples straddling the left border
    wt = rnd(1, c, 3, 3, seed=3, scale=0.3)
    b = rnd(1, seed=4)
    y, proj = ops.deform1_conv_fwd(x, off, wt, b)
    xs, os_, ws, bs = (t.double().cpu().requires_grad_(True) for t in (x, off, wt, b))
    ref = O.deformable_conv2d(xs, os_, ws, bs)
    assert y.shape == (n, 1, h, w) and rel_l2(y, ref) < 2e-6
    dy = rnd(n, 1, h, w, seed=5)
    gx, goff, gw, gb = torch.autograd.grad((ref * dy.double().cpu()).sum(), [xs, os_, ws, bs])
    dw, db = torch.zeros_like(wt), torch.zeros_like(b)
    dx = torch.full_like(x, float("nan"))            # written, not accumulated
    doff = ops.deform1_conv_bwd(x, off, wt, proj, dy, dw, db, dx)
    assert rel_l2(dw, gw) < 5e-6 and rel_l2(db, gb) < 5e-6
    assert rel_l2(dx, gx) < 5e-6
    assert rel_l2(doff, goff) < 5e-5
    dx2 = torch.ones_like(x)
    ops.deform1_conv_bwd(x, off, wt, proj, dy, torch.zeros_like(wt), torch.zeros_like(b), dx2, accumulate_dx=True)
    assert rel_l2(dx2 - 1, gx) < 5e-5


def test_bn_lrelu_fwd_bwd(ops):
    from oracle import deepbedmap_oracle as O
    n, c, h = 6, 10, 5
    x = rnd(n, c, h, h, seed=1) * 3 + 1
    gamma = rnd(c, seed=2) * 0.3 + 1
    beta = rnd(c, seed=3)
    am = rnd(c, seed=4)
    av = rnd(c, seed=5).abs() + 0.5
    p = {"bn/gamma": gamma.double().cpu().requires_grad_(True), "bn/beta": beta.double().cpu().requires_grad_(True),
         "bn/avg_mean": am.double().cpu(), "bn/avg_var": av.double().cpu()}
    xs = x.double().cpu().requires_grad_(True)
    for train in (True, False):
        stats = {}
        ref = F.leaky_relu(O.batch_norm(p, "bn", xs, train, stats), 0.2)
        y = ops.empty(n, c, h, h)
        mean, inv = ops.empty(c), ops.empty(c)
        am2, av2 = am.clone(), av.clone()
        ops.call("dbm_bn_lrelu_fwd_f32", x.data_ptr(), y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), am2.data_ptr(),
                 av2.data_ptr(), mean.data_ptr(), inv.data_ptr(), n, c, h * h, 1e-5, 0.9, int(train), ops.stream())
        assert rel_l2(y, ref) < 2e-6
        if train:
            assert rel_l2(am2, stats["bn/avg_mean"]) < 2e-6 and rel_l2(av2, stats["bn/avg_var"]) < 2e-6
            dy = rnd(n, c, h, h, seed=6)
            gx, gg, gb = torch.autograd.grad((ref * dy.double().cpu()).sum(), [xs, p["bn/gamma"], p["bn/beta"]])
            dx = ops.empty(n, c, h, h)
            dg, dbeta = ops.zeros(c), ops.zeros(c)
            scratch = ops.empty(2 * c)
            ops.call("dbm_bn_lrelu_bwd_f32", x.data_ptr(), y.data_ptr(), dy.data_ptr(), dx.data_ptr(), gamma.data_ptr(),
                     mean.data_ptr(), inv.data_ptr(), dg.data_ptr(), dbeta.data_ptr(), scratch.data_ptr(), n, c, h * h,
                     ops.stream())
            assert rel_l2(dx, gx) < 1e-5 and rel_l2(dg, gg) < 5e-6 and rel_l2(dbeta, gb) < 5e-6
        else:
            assert torch.equal(am2, am) and torch.equal(av2, av)


def test_losses_and_adam(ops):
    from oracle import deepbedmap_oracle as O
    from deepbedmap_b200 import train as T
    n = 6
    real = rnd(n, 1, seed=1)
    fake = rnd(n, 1, seed=2)
    rs, fs = real.double().cpu().requires_grad_(True), fake.double().cpu().requires_grad_(True)
    ones, zeros = torch.ones(n, 1, dtype=torch.int64), torch.zeros(n, 1, dtype=torch.int64)
    ref = O.calculate_discriminator_loss(rs, fs, ones, zeros)
    gr, gf = torch.autograd.grad(ref, [rs, fs])
    out, d_real, d_fake = T._ragan(real, fake, 1.0, 0.0, want_grads=True)
    assert abs(float(out[0]) - float(ref)) < 1e-6
    acc = O.binary_accuracy(torch.cat([rs, fs]).detach(), torch.cat([ones, zeros]))
    assert abs(float(out[1]) - float(acc)) < 1e-6
    assert rel_l2(d_real, gr) < 1e-5 and rel_l2(d_fake, gf) < 1e-5
    # KAT of the reference doctest (srgan_train.py:985-991) through the CUDA kernel
    out, _, _ = T._ragan(torch.tensor([[1.1], [-0.5]]).cuda(), torch.tensor([[-0.3], [1.0]]).cuda(), 1.0, 0.0, False)
    assert abs(float(out[0]) - 1.56670504) < 1e-6
    # image losses
    yp = rnd(n, 1, 36, 36, seed=3) * 0.5 + 1
    yt = rnd(n, 1, 36, 36, seed=4) * 0.5 + 1
    xt = rnd(n, 1, 9, 9, seed=5) * 0.5 + 1
    yps = yp.double().cpu().requires_grad_(True)
    content = (yps - yt.double().cpu()).abs().mean()
    topo = (F.avg_pool2d(yps, 4) - xt.double().cpu()).abs().mean()
    ss = O.ssim(yps, yt.double().cpu())
    total = 1e-2 * content + 2e-3 * topo + 5.25 * (1 - ss)
    (gy,) = torch.autograd.grad(total, [yps])
    sums, dy = ops.empty(4), ops.empty(n, 1, 36, 36)
    ops.call("dbm_gen_image_loss_f32", yp.data_ptr(), yt.data_ptr(), xt.data_ptr(), n, 36, 36, 1e-2, 2e-3, 5.25,
             sums.data_ptr(), dy.data_ptr(), ops.stream())
    s = sums.cpu().double()
    assert abs(float(s[0]) / (n * 1296) - float(content)) < 1e-5
    assert abs(float(s[1]) / (n * 81) - float(topo)) < 1e-5
    assert abs(float(s[2]) / (n * 784) - float(ss)) < 2e-5
    assert abs(float(s[3]) / (n * 1296) - float(((yps - yt.double().cpu()) ** 2).mean())) < 1e-5
    assert rel_l2(dy, gy) < 2e-4
    # Adam (Chainer variant), 3 steps
    p0 = rnd(1000, seed=6)
    grads = [rnd(1000, seed=10 + i) for i in range(3)]
    pr = {"w": p0.double().cpu().clone()}
    opt = O.ChainerAdam(alpha=1e-3, eps=1e-8)
    p, m, v = p0.clone(), ops.zeros(1000), ops.zeros(1000)
    for t, g in enumerate(grads, 1):
        opt.update(pr, {"w": g.double().cpu()})
        ops.call("dbm_adam_step_f32", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), 1000, 1e-3, 0.9, 0.999,
                 1e-8, t, 1.0, ops.stream())
    assert rel_l2(p, pr["w"]) < 1e-6


def test_error_conventions(ops):
    x = rnd(1, 3, 5, 5)
    with pytest.raises(ValueError):  # unsupported (ksize, stride)
        ops.conv2d_fwd(x, 0, 3, rnd(4, 3, 5, 5), None, ops.empty(1, 4, 1, 1), 0, 5, 1, 0)
    with pytest.raises(ValueError):  # Cin not a multiple of 32 on the tensor-core path
        ops.pack_conv3x3(rnd(32, 24, 3, 3), 32)
