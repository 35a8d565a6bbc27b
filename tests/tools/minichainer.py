"""A minimal stand-in for the parts of Chainer 7 that weiji14/deepbedmap's hot path touches, on torch-CPU float64.

WHY: the reference's arithmetic lives in Chainer 7 / CuPy / ssim-chainer, none of which is installable in this image
(SURVEY 8c), so the reference itself cannot run here. What CAN run is the reference's OWN model / loss / step code
(srgan_train.py's classes and functions, executed unmodified from /root/reference by
tests/golden/make_reference_golden.py) on top of this shim. That pins the GRAPH -- which layer feeds which, concat
order, residual scaling, the detach / eval-BatchNorm / ones quirks of the step functions, the parameter key layout --
to the reference's code instead of to a restatement of it; the PRIMITIVES below (convolution, deformable
convolution, batch normalisation, losses, Adam) remain restatements of Chainer's published semantics (SURVEY
App. B), each cross-checked elsewhere (tests/test_oracle_kat.py). Test infrastructure only: nothing here is imported
by the product.

Only what the reference calls is implemented; anything else raises AttributeError loudly.
"""
from __future__ import annotations

import contextlib
import copy
import math
import types
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as TF

DT = torch.float64


# ------------------------------------------------------------------------------------------------------------
# configuration
# ------------------------------------------------------------------------------------------------------------
class _Config:
    train = True
    enable_backprop = True
    cudnn_deterministic = True


global_config = _Config()
config = global_config


@contextlib.contextmanager
def using_config(name, value):
    old = getattr(global_config, name)
    setattr(global_config, name, value)
    try:
        yield
    finally:
        setattr(global_config, name, old)


# ------------------------------------------------------------------------------------------------------------
# Variable / Parameter
# ------------------------------------------------------------------------------------------------------------
def _t(x):
    """numpy / Variable / torch -> torch float64 tensor (graph kept for Variables)."""
    if isinstance(x, Variable):
        return x.t
    if isinstance(x, torch.Tensor):
        return x.to(DT)
    return torch.as_tensor(np.asarray(x)).to(DT)


class Variable:
    __array_priority__ = 200   # numpy defers `ndarray - Variable` to Variable.__rsub__, as chainer.Variable arranges

    def __init__(self, data=None, _tensor=None):
        self.t = _tensor if _tensor is not None else _t(data)

    @property
    def array(self):
        return self.t.detach().numpy()

    data = array

    @property
    def shape(self):
        return tuple(self.t.shape)

    def __len__(self):
        return self.t.shape[0]

    def backward(self):
        self.t.backward()

    def _bin(self, other, fn):
        return Variable(_tensor=fn(self.t, _t(other) if not isinstance(other, (int, float)) else other))

    def __mul__(self, o):
        return self._bin(o, lambda a, b: a * b)

    __rmul__ = __mul__

    def __add__(self, o):
        return self._bin(o, lambda a, b: a + b)

    __radd__ = __add__

    def __sub__(self, o):
        return self._bin(o, lambda a, b: a - b)

    def __rsub__(self, o):
        return self._bin(o, lambda a, b: b - a)

    def __truediv__(self, o):
        return self._bin(o, lambda a, b: a / b)

    def __neg__(self):
        return Variable(_tensor=-self.t)


def _wrap(t):
    if not global_config.enable_backprop:
        t = t.detach()
    return Variable(_tensor=t)


class Parameter(Variable):
    def __init__(self, shape=None, initializer=None):
        self.initializer = initializer
        self.t = None
        if shape is not None:
            self.initialize(shape)

    def initialize(self, shape):
        arr = np.zeros(shape, np.float64)
        if callable(self.initializer):
            self.initializer(arr)
        elif self.initializer is not None:
            arr[...] = self.initializer
        self.t = torch.as_tensor(arr).to(DT).requires_grad_(True)

    @property
    def grad(self):
        return None if self.t is None or self.t.grad is None else self.t.grad.numpy()

    def set(self, value):
        v = torch.as_tensor(np.asarray(value)).to(DT)
        if self.t is not None and tuple(self.t.shape) != tuple(v.shape):
            raise ValueError(f"shape {tuple(v.shape)} != {tuple(self.t.shape)}")
        self.t = v.clone().requires_grad_(True)


# ------------------------------------------------------------------------------------------------------------
# initializers
# ------------------------------------------------------------------------------------------------------------
class _HeNormal:
    """chainer.initializers.HeNormal(scale, fan_option='fan_in'): N(0, (scale * sqrt(2 / fan_in))^2)."""
    rng = np.random.RandomState(0)

    def __init__(self, scale=1.0, dtype=None, fan_option="fan_in"):
        assert fan_option == "fan_in"
        self.scale = scale

    def __call__(self, array):
        fan_in = int(np.prod(array.shape[1:]))
        array[...] = self.rng.normal(0.0, self.scale * math.sqrt(2.0 / fan_in), size=array.shape)


initializers = types.SimpleNamespace(HeNormal=_HeNormal)


# ------------------------------------------------------------------------------------------------------------
# Link / Chain / Sequential
# ------------------------------------------------------------------------------------------------------------
class Link:
    def __init__(self):
        object.__setattr__(self, "_params", [])
        object.__setattr__(self, "_persistent", [])
        object.__setattr__(self, "_children", [])
        object.__setattr__(self, "_within_init_scope", False)
        self.name = None

    @contextlib.contextmanager
    def init_scope(self):
        object.__setattr__(self, "_within_init_scope", True)
        try:
            yield
        finally:
            object.__setattr__(self, "_within_init_scope", False)

    def __setattr__(self, name, value):
        if getattr(self, "_within_init_scope", False):
            if isinstance(value, Link):
                value.name = name
                self._children.append(name)
            elif isinstance(value, Parameter):
                self._params.append(name)
        object.__setattr__(self, name, value)

    def add_persistent(self, name, value):
        self._persistent.append(name)
        object.__setattr__(self, name, value)

    def __call__(self, *a, **kw):
        return self.forward(*a, **kw)

    # chainer iterates sorted(self._params) then sorted(self._children)
    def namedparams(self, include_uninit=True):
        for n in sorted(self._params):
            yield "/" + n, getattr(self, n)
        for c in sorted(self._children):
            for path, p in getattr(self, c).namedparams(include_uninit):
                yield "/" + c + path, p

    def params(self, include_uninit=True):
        for _, p in self.namedparams(include_uninit):
            yield p

    def namedpersistents(self):
        for n in self._persistent:
            yield "/" + n, getattr(self, n)
        for c in sorted(self._children):
            for path, p in getattr(self, c).namedpersistents():
                yield "/" + c + path, p

    def count_params(self):
        return sum(int(p.t.numel()) for p in self.params() if p.t is not None)

    def cleargrads(self):
        for p in self.params():
            if p.t is not None:
                p.t.grad = None

    def to_gpu(self, device=None):
        return self

    @property
    def xp(self):
        return np

    def repeat(self, n_repeat, mode="init"):
        """Link.repeat: n copies with independently re-initialised parameters, as a Sequential named '0'..'n-1'."""
        assert mode == "init"
        copies = []
        for _ in range(n_repeat):
            c = copy.deepcopy(self)
            for p in c.params():
                if p.t is not None:
                    p.initialize(tuple(p.t.shape))
            copies.append(c)
        return Sequential(*copies)


Chain = Link


class Sequential(Link):
    def __init__(self, *layers):
        super().__init__()
        with self.init_scope():
            for i, l in enumerate(layers):
                setattr(self, str(i), l)
        self._n = len(layers)

    def forward(self, x):
        for i in range(self._n):
            x = getattr(self, str(i))(x)
        return x


# ------------------------------------------------------------------------------------------------------------
# links
# ------------------------------------------------------------------------------------------------------------
def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class Convolution2D(Link):
    """L.Convolution2D(in_channels, out_channels, ksize, stride=1, pad=0, nobias=False, initialW=None): cross-correlation,
    W (out, in, kh, kw), b zeros unless nobias; in_channels=None -> initialised at the first forward."""

    def __init__(self, in_channels, out_channels, ksize=None, stride=1, pad=0, nobias=False, initialW=None,
                 initial_bias=None):
        super().__init__()
        self.out_channels, self.ksize, self.stride, self.pad = out_channels, _pair(ksize), _pair(stride), _pair(pad)
        with self.init_scope():
            self.W = Parameter(initializer=initialW)
            self.b = None if nobias else Parameter((out_channels,), initializer=0.0)
        if self.b is None:
            self._params.remove("b") if "b" in self._params else None
        if in_channels is not None:
            self.W.initialize((out_channels, in_channels) + self.ksize)

    def forward(self, x):
        x = _t(x)
        if self.W.t is None:
            self.W.initialize((self.out_channels, x.shape[1]) + self.ksize)
        return _wrap(TF.conv2d(x, self.W.t, None if self.b is None else self.b.t, stride=self.stride, padding=self.pad))


class Linear(Link):
    def __init__(self, in_size, out_size=None, nobias=False, initialW=None, initial_bias=None):
        super().__init__()
        self.out_size = out_size
        with self.init_scope():
            self.W = Parameter(initializer=initialW)
            self.b = Parameter((out_size,), initializer=0.0)
        if in_size is not None:
            self.W.initialize((out_size, in_size))

    def forward(self, x):
        x = _t(x)
        x = x.reshape(x.shape[0], -1)
        if self.W.t is None:
            self.W.initialize((self.out_size, x.shape[1]))
        return _wrap(TF.linear(x, self.W.t, self.b.t))


class BatchNormalization(Link):
    """L.BatchNormalization(axis=(0, 2, 3), eps): gamma = 1, beta = 0 created at the first forward; decay 0.9; train: batch
    mean / biased variance, running variance updated with the unbiased estimate; eval: running statistics."""

    def __init__(self, size=None, decay=0.9, eps=2e-5, axis=None):
        super().__init__()
        assert axis == (0, 2, 3)
        self.decay, self.eps = decay, eps
        with self.init_scope():
            self.gamma = Parameter(initializer=1.0)
            self.beta = Parameter(initializer=0.0)
        self.add_persistent("avg_mean", None)
        self.add_persistent("avg_var", None)
        self.add_persistent("N", 0)

    def forward(self, x):
        x = _t(x)
        c = x.shape[1]
        if self.gamma.t is None:
            self.gamma.initialize((c,))
            self.beta.initialize((c,))
            self.avg_mean = np.zeros(c)
            self.avg_var = np.ones(c)
        g, b = self.gamma.t.view(1, c, 1, 1), self.beta.t.view(1, c, 1, 1)
        if global_config.train:
            mean = x.mean(dim=(0, 2, 3))
            var = x.var(dim=(0, 2, 3), unbiased=False)
            m = x.numel() // c
            self.avg_mean = self.decay * self.avg_mean + (1 - self.decay) * mean.detach().numpy()
            self.avg_var = self.decay * self.avg_var + (1 - self.decay) * var.detach().numpy() * m / max(m - 1, 1)
        else:
            mean, var = torch.as_tensor(self.avg_mean), torch.as_tensor(self.avg_var)
        y = (x - mean.view(1, c, 1, 1)) / torch.sqrt(var.view(1, c, 1, 1) + self.eps) * g + b
        return _wrap(y)


def deformable_convolution_2d_sampler(x, offset, W, b):
    """F.deformable_convolution_2d_sampler(ksize 3, stride 1, pad 1): offset channels [0:9] = dx, [9:18] = dy of tap
    t = ky*3+kx; bilinear sampling of the zero-padded input, zero outside."""
    N, C, H, Wd = x.shape
    O = W.shape[0]
    ys = torch.arange(H, dtype=DT).view(1, 1, H, 1)
    xs = torch.arange(Wd, dtype=DT).view(1, 1, 1, Wd)
    kx = torch.tensor([0, 1, 2] * 3, dtype=DT).view(1, 9, 1, 1)
    ky = torch.tensor([0, 0, 0, 1, 1, 1, 2, 2, 2], dtype=DT).view(1, 9, 1, 1)
    px = (xs + kx - 1.0 + offset[:, :9]).clamp(-2.0, Wd + 1.0)
    py = (ys + ky - 1.0 + offset[:, 9:]).clamp(-2.0, H + 1.0)
    x0, y0 = torch.floor(px), torch.floor(py)
    fx, fy = px - x0, py - y0
    x0, y0 = x0.long(), y0.long()
    flat = x.reshape(N, C, H * Wd)

    def corner(yi, xi):
        valid = ((yi >= 0) & (yi < H) & (xi >= 0) & (xi < Wd)).to(DT)
        idx = (yi.clamp(0, H - 1) * Wd + xi.clamp(0, Wd - 1)).view(N, 1, -1).expand(N, C, -1)
        return torch.gather(flat, 2, idx).view(N, C, 9, H, Wd) * valid.unsqueeze(1)

    s = (corner(y0, x0) * ((1 - fy) * (1 - fx)).unsqueeze(1) + corner(y0, x0 + 1) * ((1 - fy) * fx).unsqueeze(1)
         + corner(y0 + 1, x0) * (fy * (1 - fx)).unsqueeze(1) + corner(y0 + 1, x0 + 1) * (fy * fx).unsqueeze(1))
    return torch.einsum("nckhw,ock->nohw", s, W.reshape(O, C, 9)) + b.view(1, O, 1, 1)


class _DeformSampler(Link):
    def __init__(self, out_channels, initialW):
        super().__init__()
        self.out_channels = out_channels
        with self.init_scope():
            self.W = Parameter(initializer=initialW)
            self.b = Parameter((out_channels,), initializer=0.0)


class DeformableConvolution2D(Link):
    """L.DeformableConvolution2D = offset_conv (Convolution2D in -> 2*kh*kw, same k/s/p, with bias) + deform_conv."""

    def __init__(self, in_channels, out_channels, ksize, stride=1, pad=0, offset_nobias=False, offset_initialW=None,
                 offset_initial_bias=None, deform_nobias=False, deform_initialW=None, deform_initial_bias=None):
        super().__init__()
        assert _pair(ksize) == (3, 3) and _pair(stride) == (1, 1) and _pair(pad) == (1, 1)
        with self.init_scope():
            self.offset_conv = Convolution2D(in_channels, 18, ksize, stride, pad, initialW=offset_initialW)
            self.deform_conv = _DeformSampler(out_channels, deform_initialW)

    def forward(self, x):
        x = _t(x)
        if self.deform_conv.W.t is None:
            self.deform_conv.W.initialize((self.deform_conv.out_channels, x.shape[1], 3, 3))
        off = self.offset_conv(x).t
        return _wrap(deformable_convolution_2d_sampler(x, off, self.deform_conv.W.t, self.deform_conv.b.t))


links = types.SimpleNamespace(Convolution2D=Convolution2D, Linear=Linear, BatchNormalization=BatchNormalization,
                              DeformableConvolution2D=DeformableConvolution2D)


# ------------------------------------------------------------------------------------------------------------
# functions
# ------------------------------------------------------------------------------------------------------------
def _f_concat(xs, axis=1):
    return _wrap(torch.cat([_t(x) for x in xs], dim=axis))


def _f_leaky_relu(x, slope=0.2):
    return _wrap(TF.leaky_relu(_t(x), slope))


def _f_add(*xs):
    out = _t(xs[0])
    for x in xs[1:]:
        out = out + _t(x)
    return _wrap(out)


def _f_resize_images(x, output_shape, mode="bilinear", align_corners=True):
    assert mode == "nearest"
    x = _t(x)
    H, W = x.shape[-2:]
    oh, ow = output_shape
    iy = torch.clamp((torch.arange(oh) * H) // oh, max=H - 1)     # floor(i * H / out_H)
    ix = torch.clamp((torch.arange(ow) * W) // ow, max=W - 1)
    return _wrap(x[:, :, iy][:, :, :, ix])


def _f_reshape(x, shape):
    return _wrap(_t(x).reshape(shape))


def _f_mean_absolute_error(x0, x1):
    return _wrap((_t(x0) - _t(x1)).abs().mean())


def _f_average_pooling_2d(x, ksize, stride=None, pad=0):
    k = _pair(ksize)
    return _wrap(TF.avg_pool2d(_t(x), k, stride=k if stride is None else _pair(stride), padding=pad))


def _f_mean(x):
    return _wrap(_t(x).mean())


def _f_sigmoid_cross_entropy(x, t, normalize=True, reduce="mean"):
    x = _t(x)
    t = torch.as_tensor(np.asarray(t.array if isinstance(t, Variable) else t))
    mask = (t != -1)
    tt = t.to(DT)
    loss = -(x * (tt - (x >= 0).to(DT)) - torch.log1p(torch.exp(-x.abs()))) * mask.to(DT)
    count = max(int(mask.sum()), 1) if normalize else max(x.shape[0], 1)
    return _wrap(loss.sum() / count)


def _f_binary_accuracy(y, t):
    y = _t(y)
    t = torch.as_tensor(np.asarray(t))
    return _wrap(((y >= 0).to(torch.int64) == t.to(torch.int64)).to(DT).mean())


functions = types.SimpleNamespace(concat=_f_concat, leaky_relu=_f_leaky_relu, add=_f_add, resize_images=_f_resize_images,
                                  reshape=_f_reshape, mean_absolute_error=_f_mean_absolute_error,
                                  average_pooling_2d=_f_average_pooling_2d, mean=_f_mean,
                                  sigmoid_cross_entropy=_f_sigmoid_cross_entropy, binary_accuracy=_f_binary_accuracy)


# ------------------------------------------------------------------------------------------------------------
# ssim-chainer (ssim.functions.ssim_loss): pytorch-ssim port, Gaussian window sigma 1.5, valid padding, global mean
# ------------------------------------------------------------------------------------------------------------
def _ssim_loss(y, t, window_size=11, stride=1):
    y, t = _t(y), _t(t)
    c = y.shape[1]
    g = torch.tensor([math.exp(-((i - window_size // 2) ** 2) / (2 * 1.5 ** 2)) for i in range(window_size)], dtype=DT)
    g = g / g.sum()
    w = (g[:, None] * g[None, :]).expand(c, 1, window_size, window_size).contiguous()
    conv = lambda a: TF.conv2d(a, w, stride=stride, groups=c)
    mu1, mu2 = conv(y), conv(t)
    s1, s2, s12 = conv(y * y) - mu1 * mu1, conv(t * t) - mu2 * mu2, conv(y * t) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return _wrap((((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))).mean())


ssim_module = types.SimpleNamespace(functions=types.SimpleNamespace(ssim_loss=_ssim_loss))


# ------------------------------------------------------------------------------------------------------------
# optimizer
# ------------------------------------------------------------------------------------------------------------
class Adam:
    """chainer.optimizers.Adam: m += (1-b1)(g-m); v += (1-b2)(g^2-v); p -= alpha*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps)."""

    def __init__(self, alpha=0.001, beta1=0.9, beta2=0.999, eps=1e-8):
        self.alpha, self.beta1, self.beta2, self.eps, self.t = alpha, beta1, beta2, eps, 0
        self.state = {}

    def setup(self, link):
        self.target = link
        return self

    def update(self):
        self.t += 1
        lr = self.alpha * math.sqrt(1.0 - self.beta2 ** self.t) / (1.0 - self.beta1 ** self.t)
        with torch.no_grad():
            for path, p in self.target.namedparams():
                if p.t is None or p.t.grad is None:
                    continue
                m, v = self.state.setdefault(path, (torch.zeros_like(p.t), torch.zeros_like(p.t)))
                m += (1 - self.beta1) * (p.t.grad - m)
                v += (1 - self.beta2) * (p.t.grad * p.t.grad - v)
                p.t -= lr * m / (torch.sqrt(v) + self.eps)


optimizers = types.SimpleNamespace(Adam=Adam)


# ------------------------------------------------------------------------------------------------------------
# serializers / backend
# ------------------------------------------------------------------------------------------------------------
def state_dict(link) -> "OrderedDict[str, np.ndarray]":
    """What chainer.serializers.save_npz(file, link) stores: parameters and persistents keyed by link path."""
    out = OrderedDict()
    for path, p in link.namedparams():
        out[path[1:]] = None if p.t is None else p.t.detach().numpy().copy()
    for path, v in link.namedpersistents():
        out[path[1:]] = None if v is None else np.array(v)
    return out


def load_state(link, values):
    """chainer.serializers.load_npz: assign every stored array to the parameter / persistent of the same path."""
    params = dict(link.namedparams())
    for k, v in values.items():
        if "/" + k in params:
            params["/" + k].set(v)
            continue
        parts = k.split("/")
        obj = link
        for p in parts[:-1]:
            obj = getattr(obj, p)
        if parts[-1] not in obj._persistent:
            raise KeyError(k)
        object.__setattr__(obj, parts[-1], np.asarray(v, np.float64) if np.ndim(v) else int(v))


serializers = types.SimpleNamespace(save_npz=lambda file, obj, compression=True: np.savez(file, **state_dict(obj)),
                                    load_npz=lambda file, obj: load_state(obj, dict(np.load(file))))
backend = types.SimpleNamespace(get_array_module=lambda *a: np)
variable = types.SimpleNamespace(Variable=Variable)


def as_module():
    """A module object that can stand in for ``import chainer`` in an exec namespace."""
    m = types.ModuleType("chainer")
    for k, v in dict(Chain=Chain, Link=Link, Sequential=Sequential, Variable=Variable, Parameter=Parameter,
                     initializers=initializers, links=links, functions=functions, optimizers=optimizers,
                     serializers=serializers, backend=backend, variable=variable, global_config=global_config,
                     config=config, using_config=using_config).items():
        setattr(m, k, v)
    return m


def load_reference_namespace(path: str, names):
    """Execute the class / function definitions ``names`` of a reference notebook script (jupytext .py) -- and nothing
    else of it -- in a namespace where ``chainer``, ``F``, ``L``, ``cupy``, ``ssim`` resolve to this shim (the same
    filter the reference's own notebook loader applies, features/environment.py:15-51: keep definitions, drop
    statements)."""
    import ast
    import typing
    src = open(path).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in names]
    missing = set(names) - {n.name for n in keep}
    if missing:
        raise KeyError(f"not defined in {path}: {sorted(missing)}")
    cupy = types.SimpleNamespace(ndarray=np.ndarray, is_available=lambda: False, asnumpy=np.asarray)
    ns = {"chainer": as_module(), "F": functions, "L": links, "np": np, "cupy": cupy, "ssim": ssim_module,
          "typing": typing, "__name__": "reference_srgan_train"}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns
