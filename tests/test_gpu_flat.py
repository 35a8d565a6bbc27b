"""GPU parity of the tensor-core TRAINING trunk (csrc/umma_flat.cu, deepbedmap_b200/flat.py).

Kernel level: forward / data-gradient / weight-gradient tcgen05 GEMMs against fp64 torch convolutions
(and their autograd) of the SAME bf16-rounded operands: relative L2 <= 1e-5 for fp32 outputs, <= 4e-3 for
bf16 outputs (one storage rounding).
Trunk / model level: forward, d/d(input) and every parameter gradient with train_precision="bf16" (bf16
operands, fp32 accumulation, fp32 residual stream and gradient accumulation) against autograd on the fp64 oracle
 (a) carrying the SAME stated operand rounding (oracle.trunk_forward(emulate_bf16=True)): forward <= 2e-3
     (accumulation order flips a bf16 storage rounding here and there; measured 2e-4 ... 7e-4), d/d(a0) <= 1e-2
     (measured 3e-3 ... 4e-3), parameter gradients: median <= 1.5e-2 (measured 5e-3), worst <= 1e-1 (measured
     2e-2 ... 8e-2 on these 243 ... 405-pixel batches: the layers behind a LeakyReLU, see (b));
 (b) exact: forward <= 1e-2 / 2e-2; gradients <= 2.5e-1 / 3e-1 worst -- LeakyReLU sign flips: ANY forward with
     relative error e flips ~e of the pre-activation signs and changes those gradient elements 5x, i.e.
     ~0.8 sqrt(e) in relative L2 (4e-2 for bf16 operands; the same mechanism is the 2e-3 of the fp32 path).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a = a.double().cpu()
    b = b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def bf(t):
    return t.to(torch.bfloat16).to(torch.float64).cpu()


@pytest.fixture(scope="module")
def flat():
    from deepbedmap_b200 import flat as f
    return f


@pytest.mark.parametrize("n,h,w,cin,cout", [(3, 9, 9, 96, 32), (2, 9, 9, 192, 64), (5, 7, 12, 128, 64), (1, 20, 33, 64, 32)])
def test_flat_conv_forward(flat, n, h, w, cin, cout):
    from deepbedmap_b200 import ops
    geom = flat.geometry(n, h, w)
    assert geom == flat.geometry_host(n, h, w)
    x = rnd(n, cin, h, w, seed=1)
    wt = rnd(cout, cin, 3, 3, seed=2, scale=0.05)
    bias = rnd(cout, seed=3)
    xin = flat.alloc_bf16(cin, geom)
    flat.from_nchw(x, dst8=xin)
    assert rel_l2(flat.to_nchw(xin, cin, n, h, w), bf(x)) == 0.0
    wq = ops.pack_conv3x3(wt, cout, ck=16)
    of = flat.alloc_f32(cout, geom)
    ob = flat.alloc_bf16(cout, geom)
    pg = geom["Pg"]
    blocks = [dict(bias=bias.data_ptr() + 128 * b, act=1, out_f32=of.data_ptr() + 4 * 32 * b * pg,
                   out_bf16=ob.data_ptr() + 2 * 32 * b * pg) for b in range(cout // 32)]
    flat.conv3x3(xin, cin, wq, cout, blocks, n, h, w)
    ref = F.leaky_relu(F.conv2d(bf(x), bf(wt), bias.double().cpu(), padding=1), 0.2)
    assert rel_l2(flat.to_nchw(of, cout, n, h, w), ref) < 1e-5
    assert rel_l2(flat.to_nchw(ob, cout, n, h, w), ref) < 4e-3
    # border and guard positions are never written: they must still be exactly zero
    total = float(ob.float().abs().sum())
    inner = float(flat.to_nchw(ob, cout, n, h, w).abs().sum())
    assert abs(total - inner) <= 1e-3 * max(1.0, inner)


def test_flat_conv_dgrad_and_epilogue(flat):
    """N = 192 data-gradient GEMM (conv5 of a dense block) with every epilogue feature."""
    from deepbedmap_b200 import ops
    n, h, w, cin, cout = 3, 9, 9, 192, 64
    geom = flat.geometry(n, h, w)
    pg = geom["Pg"]
    wt = rnd(cout, cin, 3, 3, seed=2, scale=0.05)
    g = rnd(n, cout, h, w, seed=4)
    a1 = rnd(n, 64, h, w, seed=5)
    a2 = rnd(n, 32, h, w, seed=6)
    act = rnd(n, 32, h, w, seed=7)          # forward activation whose sign gates the top block
    gin = flat.alloc_bf16(cout, geom)
    flat.from_nchw(g, dst8=gin)
    a1f, a2f, actb = flat.alloc_f32(64, geom), flat.alloc_f32(32, geom), flat.alloc_bf16(32, geom)
    flat.from_nchw(a1, dst4=a1f)
    flat.from_nchw(a2, dst4=a2f)
    flat.from_nchw(act, dst8=actb)
    wq = flat.pack_dgrad(wt)
    of = flat.alloc_f32(192, geom)
    ob = flat.alloc_bf16(32, geom)
    blocks = []
    for b in range(6):
        kw = {}
        if b < 2:
            kw.update(add1=a1f.data_ptr() + 4 * 32 * b * pg, s1=0.5, beta=1.0)
        if b == 2:
            kw.update(add2=a2f.data_ptr(), beta2=0.25)
        if b < 5:
            kw.update(out_f32=of.data_ptr() + 4 * 32 * b * pg)
        else:
            kw.update(mask=actb.data_ptr(), out_bf16=ob.data_ptr(), out_scale=2.0)
        blocks.append(kw)
    flat.conv3x3(gin, cout, wq, 192, blocks, n, h, w)
    xz = torch.zeros(n, cin, h, w, dtype=torch.float64, requires_grad=True)
    F.conv2d(xz, bf(wt), padding=1).backward(bf(g))
    ref = xz.grad.clone()
    ref[:, :64] += 0.5 * a1.double().cpu()
    ref[:, 64:96] = a2.double().cpu() + 0.25 * ref[:, 64:96]
    top = ref[:, 160:] * torch.where(bf(act) >= 0, 1.0, 0.2) * 2.0
    got = flat.to_nchw(of, 160, n, h, w)
    assert rel_l2(got, ref[:, :160]) < 1e-5
    assert rel_l2(flat.to_nchw(ob, 32, n, h, w), top) < 4e-3


@pytest.mark.parametrize("n,h,w,cin,nsplit", [(3, 9, 9, 160, 2), (128, 9, 9, 192, 3), (2, 14, 11, 64, 1)])
def test_flat_wgrad_and_bias_grad(flat, n, h, w, cin, nsplit):
    from deepbedmap_b200 import ops
    geom = flat.geometry(n, h, w)
    pg = geom["Pg"]
    a = rnd(n, cin, h, w, seed=1)
    g = rnd(n, 32, h, w, seed=2)
    ab, gb = flat.alloc_bf16(cin, geom), flat.alloc_bf16(32, geom)
    flat.from_nchw(a, dst8=ab)
    flat.from_nchw(g, dst8=gb)
    dw = ops.zeros(32, cin, 3, 3)
    db = ops.zeros(32)
    splits = flat.split_blocks(geom["tiles"], nsplit)
    chunks = flat.chunk_channels(cin)
    units, reduces = [], []
    for c0, nch in chunks:
        first = len(units)
        for blk0, nblk in splits:
            units.append((ab.data_ptr() + 2 * c0 * pg, gb.data_ptr(), len(units), blk0, nblk, nch // 8, 0, (0, 0)))
        reduces.append((first, dw.data_ptr(), flat.PARTIAL_FLOATS, len(splits), cin, c0, 0, nch, 0))
    partial = torch.full((len(units) * flat.PARTIAL_FLOATS,), float("nan"), device="cuda")
    u = np.array(units, dtype=flat.WGRAD_UNIT_DTYPE)
    u["partial"] = partial.data_ptr() + u["partial"] * np.uint64(flat.PARTIAL_FLOATS * 4)
    rd = np.array(reduces, dtype=flat.WGRAD_REDUCE_DTYPE)
    rd["partial"] = partial.data_ptr() + rd["partial"] * np.uint64(flat.PARTIAL_FLOATS * 4)
    bg = np.array([(gb.data_ptr(), db.data_ptr())], dtype=flat.BIAS_GRAD_DTYPE)
    dev = lambda t: torch.from_numpy(t.view(np.uint8).reshape(-1).copy()).cuda()
    ud, rdd, bgd = dev(u), dev(rd), dev(bg)
    st = ops.stream()
    ops.call("dbm_flat_wgrad", ud.data_ptr(), len(u), n, h, w, st)
    ops.call("dbm_flat_wgrad_reduce", rdd.data_ptr(), len(rd), st)
    ops.call("dbm_flat_bias_grad", bgd.data_ptr(), 1, n, h, w, st)
    wz = torch.zeros(32, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(bf(a), wz, padding=1).backward(bf(g))
    assert rel_l2(dw, wz.grad) < 1e-5
    assert rel_l2(db, bf(g).sum(dim=(0, 2, 3))) < 1e-5


@pytest.mark.parametrize("nb,n,beta,ic", [(1, 3, 0.1, 32), (2, 5, 0.2, 32), (12, 2, 0.1, 32), (1, 3, 0.1, 64), (2, 4, 0.2, 64)])
def test_flat_trunk_alone_matches_oracle_autograd(nb, n, beta, ic):
    """The trunk in isolation (no deformable head in the way): a3 = a1 + post_res(RRDB^nb(a1)),
    a1 = lrelu(pre_res(a0)) forward, and d/d(a0), d/d(every trunk parameter) of sum(a3 * da3), against
    autograd on the oracle's trunk (a) with the SAME stated operand rounding (tight) and (b) exact (loose:
    LeakyReLU sign flips, see oracle.trunk_forward). ``ic`` = inter_channels (srgan_train.py:283-284): 64 is the wide
    setting of the reference's search space -- the flat layer chain, data gradients of conv4 / conv5 in two N-slices."""
    from oracle import deepbedmap_oracle as O
    from deepbedmap_b200 import GeneratorModel
    params = O.init_generator_params(nb, seed=3, bias_std=0.05, scale=1.0, inter_channels=ic)
    m = GeneratorModel(num_residual_blocks=nb, residual_scaling=beta, inter_channels=ic)
    assert m.train_precision == "bf16"
    for k, v in params.items():
        m.set_param(k, v)
    H = W = 9
    a0 = rnd(n, 128, H, W, seed=11)
    da3 = rnd(n, 64, H, W, seed=12)
    ft = m._flat_trunk(n, H, W)
    a3 = ft.forward(a0)
    m.cleargrads()
    da0 = ft.backward(da3)
    trunk_keys = [k for k in m.p if k.startswith(("residual_network", "pre_residual", "post_residual"))]
    for emulate, tol_fwd, tol_da0, tol_grad, tol_med in ((True, 2e-3, 1e-2, 1.5e-1, 4e-2), (False, 1e-2, 1e-1, 2.5e-1, 1e-1)):
        p64 = {k: torch.as_tensor(v, dtype=torch.float64).requires_grad_(True) for k, v in params.items()}
        a0r = a0.double().cpu().requires_grad_(True)
        a3r = O.trunk_forward(p64, a0r, nb, beta, emulate_bf16=emulate)
        (a3r * da3.double().cpu()).sum().backward()
        e_fwd, e_da0 = rel_l2(a3, a3r.detach()), rel_l2(da0, a0r.grad)
        errs = {k: rel_l2(m.g[k], p64[k].grad) for k in trunk_keys}
        worst = max(errs, key=errs.get)
        print(f"trunk alone nb={nb} ic={ic} vs {'bf16-operand' if emulate else 'exact'} oracle: forward {e_fwd:.2e}, "
              f"d/d(a0) {e_da0:.2e}, worst parameter gradient {worst} {errs[worst]:.2e}, "
              f"median {float(np.median(list(errs.values()))):.2e}")
        assert e_fwd < tol_fwd and e_da0 < tol_da0
        assert errs[worst] < tol_grad, (worst, errs[worst])
        assert float(np.median(list(errs.values()))) < tol_med


@pytest.mark.parametrize("nb,n,H,W,ic", [(1, 3, 9, 9, 32), (12, 128, 9, 9, 32), (2, 40, 30, 23, 32), (1, 1, 5, 4, 32),
                                         (2, 128, 9, 9, 64), (1, 7, 12, 10, 64)])
def test_persistent_chain_equals_per_layer_launches(nb, n, H, W, ic):
    """The one-launch forward / data-gradient chains (flag-synchronised tiles) are bit-identical to launching
    every layer on its own; (2, 40, 30, 23) gives several tiles per CTA (more than 148 tiles)."""
    from oracle import deepbedmap_oracle as O
    from deepbedmap_b200 import GeneratorModel
    params = O.init_generator_params(nb, seed=5, bias_std=0.05, scale=0.7, inter_channels=ic)
    m = GeneratorModel(num_residual_blocks=nb, inter_channels=ic)
    for k, v in params.items():
        m.set_param(k, v)
    a0 = rnd(n, 128, H, W, seed=21)
    da3 = rnd(n, 64, H, W, seed=22)
    ft = m._flat_trunk(n, H, W)
    ft.local = False   # the flat chain itself (small tiles default to the image-resident kernel)
    out = {}
    for persistent in (False, True, True):
        ft.persistent = persistent
        a3 = ft.forward(a0).clone()
        m.cleargrads()
        da0 = ft.backward(da3).clone()
        res = (a3, da0, m.flat_grad.clone())
        if persistent in out:
            assert all(torch.equal(a, b) for a, b in zip(out[persistent][:2], res[:2]))   # re-run: flags reset
        out[persistent] = res
    assert torch.isfinite(out[True][0]).all() and out[True][0].abs().max() > 0
    assert torch.equal(out[True][0], out[False][0]), "forward chain differs"
    assert torch.equal(out[True][1], out[False][1]), "data-gradient chain differs"
    # the weight gradients read the chains' bf16 outputs: same inputs -> same partial sums (fp32 split sums are
    # reduced in a fixed order)
    assert rel_l2(out[True][2], out[False][2]) < 1e-6


@pytest.mark.parametrize("nb,n,H,W", [(1, 3, 9, 9), (12, 128, 9, 9), (2, 5, 9, 9), (3, 301, 9, 9), (1, 2, 6, 14),
                                      (1, 1, 5, 4)])
def test_image_resident_trunk_equals_flat_chain(nb, n, H, W):
    """csrc/umma_local.cu (activations of an image pair resident in shared memory / TMEM, input-stationary
    passes of N = 192..64) against the layer-by-layer flat chain: every conv column accumulates the same
    products in the same order and the fused skip adds are the same explicit fma, so the fp32 output and
    the bf16 activations kept for backward are BIT-IDENTICAL; (3, 301) = odd batch, several image pairs
    per CTA."""
    from oracle import deepbedmap_oracle as O
    from deepbedmap_b200 import GeneratorModel
    params = O.init_generator_params(nb, seed=5, bias_std=0.05, scale=0.7)
    m = GeneratorModel(num_residual_blocks=nb)
    for k, v in params.items():
        m.set_param(k, v)
    a0 = rnd(n, 128, H, W, seed=21)
    da3 = rnd(n, 64, H, W, seed=22)
    ft = m._flat_trunk(n, H, W)
    assert ft.local_dev is not None
    res = {}
    for local in (False, True, True):
        ft.local = local
        for c in ft.cat:
            c.zero_()
        a3 = ft.forward(a0).clone()
        cats = [c.clone() for c in ft.cat]
        m.cleargrads()
        da0 = ft.backward(da3).clone()
        cur = (a3, cats, da0, m.flat_grad.clone())
        if local in res:   # second run of the image-resident kernel: bit-identical
            assert torch.equal(res[local][0], a3) and all(torch.equal(a, b) for a, b in zip(res[local][1], cats))
        res[local] = cur
    ref, got = res[False], res[True]
    assert torch.isfinite(got[0]).all() and got[0].abs().max() > 0
    assert torch.equal(got[0], ref[0]), "a3 differs"
    for j, (a, b) in enumerate(zip(got[1], ref[1])):
        assert torch.equal(a, b), f"kept activations of dense block {j} differ"
    # backward: the image-resident data-gradient chain accumulates the five convs' contributions to a slot in ONE
    # TMEM accumulator (the flat chain adds five separately rounded fp32 sums): equal up to fp32 association and the
    # rare bf16 flips it causes downstream
    e_da0, e_grad = rel_l2(got[2], ref[2]), rel_l2(got[3], ref[3])
    print(f"image-resident backward vs flat chain: d a0 rel_l2 {e_da0:.2e}, weight/bias gradients rel_l2 {e_grad:.2e}")
    assert e_da0 < 2e-3 and e_grad < 2e-3


@pytest.mark.parametrize("nb,n,scale,ic", [(1, 3, 0.5, 32), (2, 5, 0.5, 32), (12, 2, 0.5, 32), (2, 3, 0.5, 64)])
def test_generator_tensor_core_backward_matches_oracle(nb, n, scale, ic):
    """Whole-generator G-step gradients with the trunk on the tensor cores (stem and head fp32):
    against autograd on the oracle graph whose trunk carries the same operand rounding. (12, 2, 0.5) is the
    reference's full depth: all 384 parameter arrays of the 12-RRDB generator."""
    from oracle import deepbedmap_oracle as O
    from deepbedmap_b200 import GeneratorModel
    params = O.init_generator_params(nb, seed=0, bias_std=0.05, scale=scale, inter_channels=ic)
    m = GeneratorModel(num_residual_blocks=nb, residual_scaling=0.1, precision="bf16", train_precision="bf16",
                       inter_channels=ic)
    for k, v in params.items():
        m.set_param(k, v)
    ins = O.synthetic_inputs(n, 11, 11)
    dy = np.random.RandomState(7).randn(n, 1, 36, 36).astype(np.float32)
    y = m.forward_train(*ins).array.clone()
    m.cleargrads()
    m.backward(torch.as_tensor(dy).cuda())
    tins = [torch.as_tensor(a, dtype=torch.float64) for a in ins]
    for emulate, tol_fwd, tol_grad, tol_med in ((True, 5e-3, 2e-1, 5e-2), (False, 2e-2, 3e-1, 1e-1)):
        p64 = {k: torch.as_tensor(v, dtype=torch.float64).requires_grad_(True) for k, v in params.items()}
        y_ref = O._generator_forward_exact(p64, *tins, num_residual_blocks=nb, trunk_bf16=emulate)
        (y_ref * torch.as_tensor(dy, dtype=torch.float64)).sum().backward()
        e_fwd = rel_l2(y, y_ref.detach())
        errs = {k: rel_l2(m.g[k], p64[k].grad) for k in m.p}
        worst = max(errs, key=errs.get)
        print(f"generator nb={nb} ic={ic} vs {'bf16-trunk' if emulate else 'exact'} oracle: forward {e_fwd:.2e}, worst "
              f"parameter gradient {worst} {errs[worst]:.2e}, median {float(np.median(list(errs.values()))):.2e}")
        assert e_fwd < tol_fwd
        assert errs[worst] < tol_grad, (worst, errs[worst])
        assert float(np.median(list(errs.values()))) < tol_med


# ---- single convolutions (discriminator, generator head) ---------------------------------------------------
FLATCONV_CASES = [  # n, C, H, W, O, k, bias, act
    (3, 64, 18, 18, 64, 3, True, True),      # generator post_upsample_conv_layer_1
    (2, 64, 36, 36, 18, 3, True, False),     # offset conv of a deformable layer (18 -> padded 32 columns)
    (4, 64, 36, 36, 64, 4, False, False),    # discriminator conv_layer1 (4x4 stride 2 -> space-to-depth)
    (5, 128, 9, 9, 256, 4, False, False),    # conv_layer5: odd input (9 -> 4), two output chunks
    (3, 256, 4, 4, 256, 3, False, False),    # conv_layer6
    (6, 512, 2, 2, 512, 4, False, False),    # conv_layer9: 2x2 -> 1x1, K = 2048 phase channels
    (2, 128, 7, 10, 128, 3, False, False),
    (3, 64, 18, 18, 128, 3, False, False),   # conv_layer2
    (3, 128, 18, 18, 128, 4, False, False),  # conv_layer3
    (3, 128, 9, 9, 128, 3, False, False),    # conv_layer4
    (5, 256, 4, 4, 512, 4, False, False),    # conv_layer7
    (5, 512, 2, 2, 512, 3, False, False),    # conv_layer8
]


@pytest.mark.parametrize("n,C,H,W,O,k,with_bias,act", FLATCONV_CASES)
def test_flat_single_conv_forward_backward(flat, n, C, H, W, O, k, with_bias, act):
    from deepbedmap_b200 import ops
    stride = 2 if k == 4 else 1
    x = rnd(n, C, H, W, seed=1)
    wt = rnd(O, C, k, k, seed=2, scale=0.05)
    bias = rnd(O, seed=3) if with_bias else None
    gw = ops.zeros(O, C, k, k)
    im = flat.ConvImages(wt, bias)
    table = flat.pack_images([im])
    fc = flat.FlatConv(im, gw, n, H, W, act=act, nslots=2)
    z = fc.forward(x, slot=1)
    xr = bf(x).requires_grad_(True)
    wr = bf(wt).requires_grad_(True)
    zr = F.conv2d(xr, wr, bias.double().cpu() if with_bias else None, stride=stride, padding=1)
    if act:
        zr = F.leaky_relu(zr, 0.2)
    assert tuple(z.shape) == tuple(zr.shape)
    assert rel_l2(z, zr.detach()) < 1e-5
    dz = rnd(*z.shape, seed=4)
    fc.forward(rnd(n, C, H, W, seed=9), slot=0)         # another pass in the other slot must not disturb slot 1
    dx = fc.backward(dz, slot=1)
    # reference gradients with the SAME bf16-rounded dz operand (the kernel rounds it when staging)
    zr2 = F.conv2d(xr, wr, None, stride=stride, padding=1)
    zr2.backward(bf(dz))
    assert rel_l2(dx, xr.grad) < 1e-5
    assert rel_l2(gw, wr.grad) < 1e-5
    del table


def test_discriminator_tensor_core_matches_oracle():
    """DiscriminatorModel(precision="bf16") in a D-step style double pass (two saved forward passes, an eval pass
    in between, two backward passes).
    Forward: logits against the exact oracle <= 3e-2 (bf16 operand rounding through nine BN + LeakyReLU layers;
    layer by layer the tensor-core convolutions match a bf16-operand fp64 convolution to 1e-5, see above, but one
    flipped bf16 rounding per few thousand activations grows to plain rounding noise within a few layers, so an
    end-to-end emulation cannot be tighter than the exact comparison).
    Backward: TEACHER-FORCED against the exact fp32 CUDA-core backward run on the SAME saved forward state (same
    BatchNorm statistics and LeakyReLU masks), so only the bf16 rounding of the backward operands separates the
    two: every parameter gradient <= 3e-2 relative L2, median <= 1e-2."""
    from oracle import deepbedmap_oracle as O
    from deepbedmap_b200 import DiscriminatorModel
    params = O.init_discriminator_params(seed=1, bias_std=0.1, scale=1.0)
    d = DiscriminatorModel(precision="bf16")
    for k in d.p:
        d.set_param(k, params[k])
    n = 16
    rs = np.random.RandomState(0)
    xr_, xf_ = rs.rand(n, 1, 36, 36).astype(np.float32), rs.rand(n, 1, 36, 36).astype(np.float32)
    gr_, gf_ = torch.as_tensor(rs.randn(n, 1).astype(np.float32)).cuda(), torch.as_tensor(rs.randn(n, 1).astype(np.float32)).cuda()
    lr_ = d.forward(xr_, train=True, save=True).array.clone()
    ctx_r = d._ctx
    lf_ = d.forward(xf_, train=True, save=True).array.clone()
    ctx_f = d._ctx
    d.forward(xf_, train=False)                     # an eval pass in between must not disturb the saved inputs
    p64 = O.to_torch(params)
    for got, x_ in ((lr_, xr_), (lf_, xf_)):
        ref = O.discriminator_forward(p64, torch.as_tensor(x_, dtype=torch.float64), train=True)
        e = rel_l2(got, ref)
        print(f"discriminator bf16 logits vs exact oracle: {e:.2e}")
        assert e < 3e-2
    grads = {}
    for mode in ("tc", "fp32"):
        d.cleargrads()
        for ctx, g_ in ((ctx_f, gf_), (ctx_r, gr_)):
            c = dict(ctx)
            if mode == "fp32":
                c["tc"] = None
            d._ctx = c
            d.backward(g_)
        grads[mode] = d.flat_grad.clone()
    errs = {k: rel_l2(grads["tc"][o:o + m], grads["fp32"][o:o + m]) for k, (o, m) in d._slices.items()}
    worst = max(errs, key=errs.get)
    med = float(np.median(list(errs.values())))
    print(f"discriminator backward, tensor cores vs fp32 kernels on the same forward state: median {med:.2e}, "
          f"worst {worst} {errs[worst]:.2e}")
    assert med < 1e-2 and errs[worst] < 3e-2, (med, worst, errs[worst])
