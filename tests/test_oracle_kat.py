"""Pins the oracle against every known answer the reference's own tests hold for the hot
path: shapes, parameter counts and the four loss / metric doctest values."""
import numpy as np
import pytest
import torch

from oracle import deepbedmap_oracle as O

T = lambda a, dt=torch.float64: torch.tensor(a, dtype=dt)


def test_param_counts():
    # srgan_train.py:446-447, 607-608
    assert O.count_params(O.generator_param_shapes()) == 8907749
    assert O.count_params(O.discriminator_param_shapes()) == 10370761
    assert len(O.generator_param_shapes()) == 384
    assert len(O.discriminator_param_shapes()) + len(O.discriminator_persistent_shapes()) == 60


def test_generator_shape_doctest():
    # srgan_train.py:437-445
    p = O.init_generator_params()
    x, w1, w2, w3 = O.synthetic_inputs(1)
    y = O.generator_forward_numpy(p, x, w1, w2, w3)
    assert y.shape == (1, 1, 36, 36)
    assert np.isfinite(y).all()


def test_generator_x4_rule():
    # features/steps/test_deepbedmap.py:38-39
    p = O.init_generator_params(num_residual_blocks=1)
    x, w1, w2, w3 = O.synthetic_inputs(1, h=14, w=9)
    y = O.generator_forward_numpy(p, x, w1, w2, w3, num_residual_blocks=1)
    assert y.shape[2] / (x.shape[2] - 2) == 4.0 and y.shape[3] / (x.shape[3] - 2) == 4.0


def test_discriminator_shape_doctest():
    # srgan_train.py:601-606
    p = O.to_torch(O.init_discriminator_params())
    x = torch.as_tensor(np.random.RandomState(0).rand(2, 1, 36, 36))
    assert tuple(O.discriminator_forward(p, x).shape) == (2, 1)


def test_discriminator_loss_kat():
    # srgan_train.py:985-991
    v = O.calculate_discriminator_loss(T([[1.1], [-0.5]]), T([[-0.3], [1.0]]),
                                       torch.tensor([[1], [1]]), torch.tensor([[0], [0]]))
    assert float(v) == pytest.approx(1.56670504, abs=5e-9)


def test_psnr_kat():
    # srgan_train.py:916-920
    v = O.psnr(torch.ones(2, 1, 3, 3, dtype=torch.float64),
               torch.full((2, 1, 3, 3), 2.0, dtype=torch.float64))
    assert float(v) == pytest.approx(192.65919722494797, rel=1e-14)


def test_ssim_kat():
    # srgan_train.py:944-948
    v = O.ssim(torch.ones(2, 1, 9, 9, dtype=torch.float64),
               torch.full((2, 1, 9, 9), 2.0, dtype=torch.float64))
    assert float(v) == pytest.approx(0.800004, abs=5e-7)
    with pytest.raises(ValueError):  # :950-951
        O.ssim(torch.ones(2, 1, 9, 9), torch.ones(2, 1, 9, 10))


def test_generator_loss_kat():
    # srgan_train.py:859-868
    v = O.calculate_generator_loss(
        y_pred=torch.ones(2, 1, 12, 12, dtype=torch.float64),
        y_true=torch.full((2, 1, 12, 12), 10.0, dtype=torch.float64),
        fake_labels=T([[-1.2], [0.5]]), real_labels=T([[0.5], [-0.8]]),
        fake_minus_real_target=torch.tensor([[1], [1]]),
        real_minus_fake_target=torch.tensor([[0], [0]]),
        x_topo=torch.full((2, 1, 3, 3), 9.0, dtype=torch.float64))
    assert float(v) == pytest.approx(4.35108415, abs=5e-9)


def _naive_deform(x, off, W, b):
    """Pure-NumPy loop restatement of SURVEY App. B.6 for a tiny case."""
    N, C, H, Wd = x.shape
    Oc = W.shape[0]
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
    y = np.zeros((N, Oc, H, Wd))

    def at(n, c, yy, xx):  # padded frame, zero outside
        if 0 <= yy < H + 2 and 0 <= xx < Wd + 2:
            return xp[n, c, yy, xx]
        return 0.0

    for n in range(N):
        for oy in range(H):
            for ox in range(Wd):
                for t in range(9):
                    ky, kx = divmod(t, 3)
                    px = ox + kx + off[n, t, oy, ox]
                    py = oy + ky + off[n, 9 + t, oy, ox]
                    x0, y0 = int(np.floor(px)), int(np.floor(py))
                    fx, fy = px - x0, py - y0
                    for c in range(C):
                        v = ((1 - fy) * (1 - fx) * at(n, c, y0, x0) + (1 - fy) * fx * at(n, c, y0, x0 + 1)
                             + fy * (1 - fx) * at(n, c, y0 + 1, x0) + fy * fx * at(n, c, y0 + 1, x0 + 1))
                        y[n, :, oy, ox] += W[:, c, ky, kx] * v
    return y + b[None, :, None, None]


def test_deformable_conv_against_naive_and_torchvision():
    rng = np.random.RandomState(3)
    x = rng.randn(2, 3, 5, 6)
    off = rng.randn(2, 18, 5, 6) * 1.7
    W = rng.randn(4, 3, 3, 3)
    b = rng.randn(4)
    got = O.deformable_conv2d(T(x), T(off), T(W), T(b)).numpy()
    np.testing.assert_allclose(got, _naive_deform(x, off, W, b), rtol=1e-10, atol=1e-10)
    # zero offsets == plain convolution
    plain = torch.nn.functional.conv2d(T(x), T(W), T(b), padding=1).numpy()
    got0 = O.deformable_conv2d(T(x), T(np.zeros_like(off)), T(W), T(b)).numpy()
    np.testing.assert_allclose(got0, plain, rtol=1e-12, atol=1e-12)
    tv = pytest.importorskip("torchvision.ops")
    # torchvision interleaves (dy, dx) per tap; Chainer stores [dx x9, dy x9]
    off_tv = np.empty_like(off)
    off_tv[:, 0::2] = off[:, 9:]
    off_tv[:, 1::2] = off[:, :9]
    ref = tv.deform_conv2d(T(x), T(off_tv), T(W), T(b), padding=1).numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-9)
    # the row-gather form bench.py times (and the full-size golden tile uses) is the same function,
    # also for offsets far outside the image
    for scale in (1.0, 6.0):
        fast = O.deformable_conv2d_fast(T(x), T(off * scale), T(W), T(b)).numpy()
        np.testing.assert_allclose(fast, O.deformable_conv2d(T(x), T(off * scale), T(W), T(b)).numpy(),
                                   rtol=1e-10, atol=1e-10)


def test_chainer_adam_formula():
    # eps is added to the UNcorrected sqrt(v)
    p = {"w": T([1.0, -2.0])}
    g = {"w": T([0.5, 0.25])}
    opt = O.ChainerAdam(alpha=1e-3, eps=1e-8)
    opt.update(p, g)
    m = 0.1 * g["w"]
    v = 0.001 * g["w"] ** 2
    lr = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    exp = T([1.0, -2.0]) - lr * m / (torch.sqrt(v) + 1e-8)
    np.testing.assert_allclose(p["w"].numpy(), exp.numpy(), rtol=1e-14)


def test_tile_plan_geometry():
    # deepbedmap.py:691-736 -- 396 tiles; interior 288x288 lowres, edge 269, placement offsets
    plan = O.tile_plan()
    assert len(plan) == 18 * 22
    shapes = {(y1 - y0, x1 - x0) for (y0, y1, x0, x1, _, _) in plan}
    assert shapes == {(288, 288), (269, 288), (288, 269), (269, 269)}
    n_int = sum(1 for (y0, y1, x0, x1, _, _) in plan if (y1 - y0, x1 - x0) == (288, 288))
    assert n_int == 320
    y0, y1, x0, x1, ys, xs = plan[0]
    assert (y0, y1, x0, x1) == (0, 269, 0, 269) and ys == slice(76, 1000) and xs == slice(76, 1000)
    y0, y1, x0, x1, ys, xs = plan[23]  # second row, second column: interior
    assert (y0, y1) == (231, 519) and ys == slice(1000, 2000)


def test_steps_change_weights_and_no_nan():
    # srgan_train.py:1100-1122, 1190-1212 (weights change) and
    # features/steps/test_srgan_train.py:60-67 (no NaN metrics), reduced to 1 RRDB
    rs = lambda *s: torch.as_tensor(np.random.RandomState(42).rand(*s))
    arrays = {"X": rs(2, 1, 11, 11), "W1": rs(2, 1, 110, 110), "W2": rs(2, 2, 22, 22),
              "W3": rs(2, 1, 11, 11), "Y": rs(2, 1, 36, 36)}
    g = O.to_torch(O.init_generator_params(1))
    d = O.to_torch(O.init_discriminator_params())
    d_w0 = d["linear_1/W"].clone()
    g_w0 = g["pre_residual_conv_layer/W"].clone()
    dl, da = O.train_eval_discriminator(arrays, g, d, O.ChainerAdam(1e-3, eps=1e-7), num_residual_blocks=1)
    gl, gp, gs = O.train_eval_generator(arrays, g, d, O.ChainerAdam(1e-3, eps=1e-7), num_residual_blocks=1)
    assert not torch.equal(d_w0, d["linear_1/W"]) and not torch.equal(g_w0, g["pre_residual_conv_layer/W"])
    assert all(np.isfinite(v) for v in (dl, da, gl, gp, gs))
    assert int(d["batch_norm1/N"]) == 2
