"""Training step with the reference's function surface (srgan_train.py:1014-1329):
compile_srgan_model, train_eval_discriminator, train_eval_generator, trainer.

Data parallelism (new; the reference is single-GPU): when torch.distributed is initialised the
flat gradient buffer of the model being trained is all-reduced (NCCL over NVLink, averaged over
ranks) in buckets that are launched from inside backward as their layers finish (last layers first),
so the exchange overlaps the remaining backward kernels; Adam runs after the last bucket. BatchNorm
statistics and the RaGAN batch means stay local to each rank, so world_size == 1 reproduces the
reference exactly.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from . import ops
from .model import DiscriminatorModel, GeneratorModel, Variable, as_device


class Adam:
    """chainer.optimizers.Adam(alpha, beta1=0.9, beta2=0.999, eps=1e-8).setup(link)
    (srgan_train.py:1043-1048); one fused kernel over the link's flat parameter buffer."""

    def __init__(self, alpha: float = 1.6e-4, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8):
        self.alpha, self.beta1, self.beta2, self.eps = alpha, beta1, beta2, eps
        self.t = 0
        self.t_dev = None
        self.target = None

    def setup(self, link):
        self.target = link
        self.m = ops.zeros(link.flat.numel())
        self.v = ops.zeros(link.flat.numel())
        return self

    def update(self, grad_scale: float = 1.0):
        link = self.target
        self.t += 1
        if self.t_dev is not None:
            # step count on the device (CUDA-graph replays, GraphedTrainStep): the kernel increments it itself
            ops.call("dbm_adam_step_dev_f32", link.flat.data_ptr(), link.flat_grad.data_ptr(), self.m.data_ptr(),
                     self.v.data_ptr(), link.flat.numel(), self.alpha, self.beta1, self.beta2, self.eps,
                     self.t_dev.data_ptr(), float(grad_scale), ops.stream())
        else:
            ops.call("dbm_adam_step_f32", link.flat.data_ptr(), link.flat_grad.data_ptr(), self.m.data_ptr(),
                     self.v.data_ptr(), link.flat.numel(), self.alpha, self.beta1, self.beta2, self.eps, self.t,
                     float(grad_scale), ops.stream())
        link.mark_updated()

    def use_device_step(self, on: bool = True):
        """Keep the step count t in device memory (needed when update() is replayed from a CUDA graph)."""
        self.t_dev = torch.full((1,), self.t, dtype=torch.int32, device="cuda") if on else None

    # -- optimizer-state checkpoint (extension: the reference saves weights only, srgan_train.py:1355-1361;
    #    SURVEY 8f N3). Keys follow chainer.serializers.save_npz(optimizer): 't', '<param path>/m', '<param path>/v'.
    def state_dict(self) -> Dict[str, np.ndarray]:
        link = self.target
        m, v = self.m.cpu().numpy(), self.v.cpu().numpy()
        out = {"t": np.asarray(self.t, np.int64), "alpha": np.asarray(self.alpha, np.float64)}
        for k, (o, n) in link._slices.items():
            out[f"{k}/m"] = m[o:o + n].reshape(link._shapes[k])
            out[f"{k}/v"] = v[o:o + n].reshape(link._shapes[k])
        return out

    def load_state_dict(self, state) -> None:
        link = self.target
        m = np.empty(link.flat.numel(), np.float32)
        v = np.empty(link.flat.numel(), np.float32)
        for k, (o, n) in link._slices.items():
            for name, buf in (("m", m), ("v", v)):
                a = np.asarray(state[f"{k}/{name}"], np.float32)
                if tuple(a.shape) != tuple(link._shapes[k]):
                    raise ValueError(f"{k}/{name}: shape {a.shape} != {link._shapes[k]}")
                buf[o:o + n] = a.reshape(-1)
        self.m.copy_(torch.from_numpy(m))
        self.v.copy_(torch.from_numpy(v))
        self.t = int(state["t"])
        if self.t_dev is not None:   # CUDA-graph replays read the step count from the device
            self.t_dev.fill_(self.t)

    def save_npz(self, file, compression: bool = True) -> None:
        (np.savez_compressed if compression else np.savez)(file, **self.state_dict())

    def load_npz(self, file) -> None:
        with np.load(file) as f:
            self.load_state_dict({k: f[k] for k in f.files})


_COMM_STREAM = None
_AUX_STREAM = None
WGRAD_SM_RESERVE = 16


def _aux_stream():
    """Side stream for work the step's critical path does not depend on (the metric-only discriminator pass)."""
    global _AUX_STREAM
    if _AUX_STREAM is None:
        # high priority: the discriminator chain is ~150 small, dependent kernels; whenever one is ready it should get
        # SMs ahead of the thousands of pending CTAs of the generator's big gather / GEMM kernels on the main stream
        _AUX_STREAM = torch.cuda.Stream(priority=-1)
    return _AUX_STREAM



def _comm_stream():
    global _COMM_STREAM
    if _COMM_STREAM is None:
        _COMM_STREAM = torch.cuda.Stream()
    return _COMM_STREAM


class GradBucketReducer:
    """Data-parallel gradient all-reduce overlapped with backward: ``bucket(lo, hi)`` (the models'
    ``on_ready`` hook) enqueues an NCCL all-reduce(SUM) of flat_grad[lo:hi] on a side stream as soon as
    backward has finished that range, so the exchange of the last layers' gradients over NVLink runs
    under the wgrad/dgrad kernels of the earlier ones. ``finish()`` makes the compute stream wait for
    all buckets and returns the 1/world_size scale the optimizer applies. With world_size == 1 (the
    reference's case) it does nothing."""

    def __init__(self, link):
        import torch.distributed as dist
        self.link = link
        self.dist = dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.works = []
        self.covered = []
        if self.dist is not None:
            # gloo (CPU tests) has no streams; NCCL buckets go on one side stream per reducer
            self.stream = _comm_stream() if link.flat_grad.is_cuda else None

    def bucket(self, lo: int, hi: int):
        self.covered.append((lo, hi))
        if self.dist is None:
            return
        view = self.link.flat_grad[lo:hi]
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())   # bucket's kernels are enqueued before this point
            with torch.cuda.stream(self.stream):
                self.works.append(self.dist.all_reduce(view, op=self.dist.ReduceOp.SUM, async_op=True))
        else:
            self.works.append(self.dist.all_reduce(view, op=self.dist.ReduceOp.SUM, async_op=True))

    def finish(self) -> float:
        total = self.link.flat_grad.numel()
        pos = 0
        for lo, hi in sorted(self.covered):
            if lo != pos:
                raise RuntimeError(f"gradient buckets do not tile the parameter buffer (gap at {pos}:{lo})")
            pos = hi
        if pos != total:
            raise RuntimeError(f"gradient buckets stop at {pos} of {total}")
        if self.dist is None:
            return 1.0
        for w in self.works:
            w.wait()                                               # current stream waits for the NCCL work
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        return 1.0 / self.dist.get_world_size()


def allreduce_grads(link) -> float:
    """Un-bucketed variant: one all-reduce of the whole flat gradient after backward."""
    r = GradBucketReducer(link)
    r.bucket(0, link.flat_grad.numel())
    return r.finish()


def set_deterministic(on: bool = True) -> None:
    """chainer.global_config.cudnn_deterministic = True (srgan_train.py:69): fixed-order gradient reductions and an
    exact fixed-point scatter in the deformable layers' backward -- two runs of the same steps give bit-identical
    weights and metrics. Process-wide; slower (the batch-reduced GEMMs run image by image)."""
    ops.call("dbm_set_deterministic", int(bool(on)))


def compile_srgan_model(num_residual_blocks: int = 12, residual_scaling: float = 0.1,
                        learning_rate: float = 1.6e-4, seed: int = 0, *, inter_channels: int = 32,
                        train_precision: Optional[str] = None):
    """srgan_train.py:1014-1055: returns (g_model, g_optimizer, d_model, d_optimizer). ``inter_channels`` is the
    dense-block growth of ResidualDenseBlock (srgan_train.py:283-284; 32 in the reference's final model, 64 the other
    value of its search space); ``train_precision`` as in GeneratorModel ("bf16" tensor-core training by default)."""
    g = GeneratorModel(num_residual_blocks=num_residual_blocks, residual_scaling=residual_scaling, seed=seed,
                       inter_channels=inter_channels, train_precision=train_precision)
    d = DiscriminatorModel(seed=seed + 1)
    # the generator's persistent weight-gradient kernel leaves a few SMs to the discriminator chain that runs beside it
    ops.call("dbm_set_sm_reserve", WGRAD_SM_RESERVE)
    g_opt = Adam(alpha=learning_rate, eps=1e-8).setup(g)
    d_opt = Adam(alpha=learning_rate, eps=1e-8).setup(d)
    return g, g_opt, d, d_opt


def _ragan(real_pred, fake_pred, t_rmf, t_fmr, want_grads, grad_scale=1.0, grads_out=None):
    n = real_pred.shape[0]
    out = ops.empty(2)
    if want_grads and grads_out is not None:      # (2n, 1): gradients wrt [real; fake] logits
        d_real, d_fake = grads_out[:n], grads_out[n:]
    else:
        d_real = ops.empty(n, 1) if want_grads else None
        d_fake = ops.empty(n, 1) if want_grads else None
    ops.call("dbm_ragan_loss_f32", real_pred.data_ptr(), fake_pred.data_ptr(), n, float(t_rmf), float(t_fmr),
             float(grad_scale), out.data_ptr(), d_real.data_ptr() if want_grads else None,
             d_fake.data_ptr() if want_grads else None, ops.stream())
    return out, d_real, d_fake


def train_eval_discriminator(input_arrays: Dict[str, object], g_model: GeneratorModel, d_model: DiscriminatorModel,
                             d_optimizer: Optional[Adam] = None, train: bool = True,
                             share_generator_forward: bool = False):
    """srgan_train.py:1084-1166 -> (d_loss, d_accu).

    ``share_generator_forward`` (used by ``trainer``): the reference runs G(x) here without a graph (:1131-1137)
    and again with one in the generator step (:1222-1227) although G's weights do not change in between. With
    the flag set, this step runs the graph-keeping forward once and ``train_eval_generator`` called next on the
    SAME device arrays reuses its output and saved activations (``GeneratorModel.shared_forward``) -- the same
    values the second forward would produce, one generator forward less per step."""
    out = _discriminator_step_enqueue(input_arrays, g_model, d_model, d_optimizer, train, share_generator_forward)
    res = out.cpu()                                                                # :1166 (host sync)
    return float(res[0]), float(res[1])


def _discriminator_step_enqueue(input_arrays, g_model, d_model, d_optimizer, train, share_generator_forward,
                                fake=None):
    """Everything train_eval_discriminator launches, without the host read: returns the device pair
    (d_loss, d_accu). ``fake``: the generator output if the caller already ran the forward."""
    if train:
        assert d_optimizer is not None  # :1127
    if fake is not None:
        pass
    elif train and share_generator_forward:
        fake = g_model.forward_train(input_arrays["X"], input_arrays["W1"], input_arrays["W2"],
                                     input_arrays["W3"]).array
    else:
        fake = g_model.forward(x=input_arrays["X"], w1=input_arrays["W1"], w2=input_arrays["W2"],
                               w3=input_arrays["W3"]).array                     # :1131-1137 (no graph)
    real = as_device(input_arrays["Y"])
    if train:
        d_model.cleargrads()                                                       # :1162
    # D(real) then D(fake), each with its own batch statistics (:1145-1146): one stacked pass, BatchNormalization
    # per group (DiscriminatorModel.forward, ``groups``)
    n = real.shape[0]
    if tuple(real.shape) != tuple(fake.shape):
        raise ValueError("Input images must have the same dimensions.")
    pred = d_model.forward(torch.cat([real, fake]), train=train, save=train, groups=2).array
    real_pred, fake_pred = pred[:n], pred[n:]
    dlogit = ops.empty(2 * n, 1) if train else None
    out, _, _ = _ragan(real_pred, fake_pred, 1.0, 0.0, want_grads=train, grads_out=dlogit)  # :1149-1158
    if train:
        reducer = GradBucketReducer(d_model)
        d_model.backward(dlogit, on_ready=reducer.bucket)                          # :1163
        d_optimizer.update(grad_scale=reducer.finish())                            # :1164
    d_model._ctx = None
    return out


def train_eval_generator(input_arrays: Dict[str, object], g_model: GeneratorModel, d_model: DiscriminatorModel,
                         g_optimizer: Optional[Adam] = None, train: bool = True,
                         content_loss_weighting: float = 1e-2, adversarial_loss_weighting: float = 2e-2,
                         topographic_loss_weighting: float = 2e-3, structural_loss_weighting: float = 5.25):
    """srgan_train.py:1170-1263 -> (g_loss, g_psnr, g_ssim)."""
    sums, adv, shape = _generator_step_enqueue(input_arrays, g_model, d_model, g_optimizer, train, content_loss_weighting,
                                               topographic_loss_weighting, structural_loss_weighting)
    return _generator_step_finalize(sums.cpu().double().numpy(), float(adv.cpu()[0]), shape, content_loss_weighting,
                                    adversarial_loss_weighting, topographic_loss_weighting, structural_loss_weighting)


def _generator_step_enqueue(input_arrays, g_model, d_model, g_optimizer, train, content_loss_weighting,
                            topographic_loss_weighting, structural_loss_weighting):
    """Everything train_eval_generator launches, without the host reads: returns the device partial sums
    (content, topographic, ssim, squared error), the device adversarial loss and the prediction shape."""
    if train:
        assert g_optimizer is not None  # :1218
    X = as_device(input_arrays["X"])
    if train:
        fake = g_model.shared_forward(X, input_arrays["W1"], input_arrays["W2"], input_arrays["W3"])
        if fake is None:
            fake = g_model.forward_train(X, input_arrays["W1"], input_arrays["W2"], input_arrays["W3"]).array
    else:
        fake = g_model.forward(X, input_arrays["W1"], input_arrays["W2"], input_arrays["W3"]).array
    real = as_device(input_arrays["Y"])
    if tuple(real.shape) != tuple(fake.shape):
        raise ValueError("Input images must have the same dimensions.")            # :950-951
    n, _, H, W = fake.shape
    # eval-mode BatchNorm and `.array`: the adversarial term carries no gradient (:1228-1229), so the generator's
    # backward does not wait for it -- when training, the discriminator pass runs on a side stream next to the image
    # losses, backward and Adam (dozens of small kernels on either side) and joins before the results are read
    cur = torch.cuda.current_stream()
    side = _aux_stream() if train else cur
    if side is not cur:
        side.wait_stream(cur)
    with torch.cuda.stream(side):
        fake_labels = d_model.forward(fake, train=False).array
        real_labels = ops.empty(n, 1)
        ops.fill(real_labels, 1.0)                                                 # :1233
        # adversarial: calculate_discriminator_loss(real=real_labels, fake=fake_labels, rmf=0, fmr=1) (:874-879)
        adv, _, _ = _ragan(real_labels, fake_labels, 0.0, 1.0, want_grads=False)
    # x_topo = X[:, :, 1:-1, 1:-1] (:1248)
    h, w = X.shape[2], X.shape[3]
    x_topo = ops.empty(n, 1, h - 2, w - 2)
    ops.call("dbm_crop_clip_f32", X.data_ptr(), h, w, x_topo.data_ptr(), n, 1, 1, h - 2, w - 2, 0, ops.stream())
    if (h - 2) * 4 != H or (w - 2) * 4 != W:
        raise ValueError("x_topo does not match the 4x4 average-pooled prediction")
    sums = ops.empty(4)
    dy = ops.empty(n, 1, H, W) if train else None
    ops.call("dbm_gen_image_loss_f32", fake.data_ptr(), real.data_ptr(), x_topo.data_ptr(), n, H, W,
             content_loss_weighting, topographic_loss_weighting, structural_loss_weighting, sums.data_ptr(),
             dy.data_ptr() if train else None, ops.stream())
    if train:
        g_model.cleargrads()                                                       # :1255
        reducer = GradBucketReducer(g_model)
        g_model.backward(dy, on_ready=reducer.bucket)                              # :1256
        g_optimizer.update(grad_scale=reducer.finish())                            # :1257
    if side is not cur:
        cur.wait_stream(side)
    return sums, adv, (n, H, W)


def _generator_step_finalize(s, adv_v, shape, content_loss_weighting, adversarial_loss_weighting,
                             topographic_loss_weighting, structural_loss_weighting):
    """Host arithmetic of the metrics (:1259-1263) from the partial sums read back from the device."""
    n, H, W = shape
    npx = n * H * W
    content = s[0] / npx
    topo = s[1] / (n * (H // 4) * (W // 4))
    ssim = s[2] / (n * (H - 8) * (W - 8))
    mse = s[3] / npx
    g_loss = (content_loss_weighting * content + adversarial_loss_weighting * adv_v
              + topographic_loss_weighting * topo + structural_loss_weighting * (1.0 - ssim))
    with np.errstate(divide="ignore"):
        g_psnr = float(20.0 * np.log10(2.0 ** 32 / np.sqrt(mse)))                  # :906-928
    return float(g_loss), g_psnr, float(ssim)


def trainer(i: int, columns: list, train_iter, dev_iter, g_model, g_optimizer, d_model, d_optimizer,
            graphed_step: Optional["GraphedTrainStep"] = None):
    """srgan_train.py:1267-1329. ``train_iter`` / ``dev_iter`` are objects with ``.epoch`` and
    ``.next()`` returning a dict of batched arrays {X, W1, W2, W3, Y}. ``graphed_step`` (extension): a
    GraphedTrainStep built on the same models / optimizers; minibatches of its batch shape are replayed from the
    captured CUDA graph, any other (e.g. a short last minibatch) takes the eager functions."""
    metrics = {mn: [] for mn in columns}
    while i == train_iter.epoch:
        # one device copy of the batch for both steps, so the generator step can reuse the forward of the first
        arrays = {k: as_device(v) for k, v in train_iter.next().items()}
        if graphed_step is not None and graphed_step.accepts(arrays):
            (dl, da), (gl, gp, gs) = graphed_step.step(arrays)
        else:
            if graphed_step is not None:
                graphed_step.refresh()     # eager updates follow: see GraphedTrainStep.refresh
            dl, da = train_eval_discriminator(arrays, g_model, d_model, d_optimizer, share_generator_forward=True)
            gl, gp, gs = train_eval_generator(arrays, g_model, d_model, g_optimizer)
            if graphed_step is not None:
                graphed_step.refresh()
        metrics["discriminator_loss"].append(dl)
        metrics["discriminator_accu"].append(da)
        metrics["generator_loss"].append(gl)
        metrics["generator_psnr"].append(gp)
        metrics["generator_ssim"].append(gs)
    while i == dev_iter.epoch:
        arrays = dev_iter.next()
        dl, da = train_eval_discriminator(arrays, g_model, d_model, train=False)
        metrics["val_discriminator_loss"].append(dl)
        metrics["val_discriminator_accu"].append(da)
        gl, gp, gs = train_eval_generator(arrays, g_model, d_model, train=False)
        metrics["val_generator_loss"].append(gl)
        metrics["val_generator_psnr"].append(gp)
        metrics["val_generator_ssim"].append(gs)
    return metrics


class GraphedTrainStep:
    """The per-minibatch body of ``trainer`` (discriminator step, then generator step on the same batch,
    srgan_train.py:1286-1308) captured ONCE as a CUDA graph and replayed per minibatch: the ~400 kernel launches
    of a step become one graph launch (the eager step leaves the GPU idle ~10 % of the time between launches).

    The batch shapes are fixed at construction; ``step(arrays)`` copies the new minibatch into the static input
    buffers, replays, and returns the same five metrics as the eager functions
    ((d_loss, d_accu), (g_loss, g_psnr, g_ssim)). Model weights, BatchNorm statistics and Adam state live in the
    same buffers as in eager mode, so eager calls (evaluation, checkpoints) can be mixed with replays.
    Data parallel: the bucketed NCCL all-reduces (GradBucketReducer, side stream) are captured with the step."""

    LOSS_WEIGHTS = dict(content_loss_weighting=1e-2, adversarial_loss_weighting=2e-2,
                        topographic_loss_weighting=2e-3, structural_loss_weighting=5.25)
    # The discriminator chain (forward, backward, update, eval-mode metric pass: ~330 small dependent kernels) is the
    # critical path of the fused step; the generator's batched trunk weight gradient runs beside the discriminator's
    # nine weight-gradient launches, and both are persistent one-CTA-per-SM kernels. Capping the generator's grid
    # shares the SMs instead of leaving the discriminator's launches the 16 reserved ones (6.89 -> 6.71 ms per step;
    # 48 / 64 / 80 / 96 / 112 CTAs: 6.99 / 6.71 / 6.78 / 6.77 / 6.90 ms).
    TRUNK_WGRAD_CTAS = 64

    def __init__(self, input_arrays: Dict[str, object], g_model, g_optimizer, d_model, d_optimizer, warmup: int = 2):
        self.g, self.g_opt, self.d, self.d_opt = g_model, g_optimizer, d_model, d_optimizer
        self.arrays = {k: as_device(v).clone() for k, v in input_arrays.items()}
        for opt in (g_optimizer, d_optimizer):
            opt.use_device_step(True)
        # Warm-up runs real steps (workspaces, packed-image plans, kernel attributes must exist before capture);
        # the training state is restored afterwards so that constructing the object does not train.
        state = self._snapshot()
        prev_ctas = getattr(self.g, "trunk_wgrad_ctas", 0)
        for _ in range(max(1, warmup)):
            self._body()
        torch.cuda.synchronize()
        self.host = torch.empty(8, dtype=torch.float32).pin_memory()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            out, sums, adv, self.shape = self._body()
            dev = torch.cat([out.reshape(-1)[:2], sums.reshape(-1)[:4], adv.reshape(-1)[:1]])
            self.host[:7].copy_(dev, non_blocking=True)
        self._restore(state)
        self.g.trunk_wgrad_ctas = prev_ctas   # eager calls outside the graph keep the whole GPU

    def _body(self):
        # Neither step's gradients depend on the other model's update: the discriminator step needs G(x) only, and
        # the generator's backward needs the image losses only (the adversarial term is detached, :1228-1229). So
        # after the one generator forward the step forks: the whole discriminator work (stacked forward, RaGAN,
        # backward, Adam, then the metric-only eval pass on the updated weights) runs on a side stream next to
        # image losses -> generator backward -> Adam on the main one, and joins before the metrics are read.
        # Same values as the sequential order of trainer() (:1286-1308).
        a = self.arrays
        self.g.trunk_wgrad_ctas = self.TRUNK_WGRAD_CTAS   # (restored by __init__ after the capture)
        fake = self.g.forward_train(a["X"], a["W1"], a["W2"], a["W3"]).array
        cur, side = torch.cuda.current_stream(), _aux_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            out = _discriminator_step_enqueue(a, self.g, self.d, self.d_opt, True, True, fake=fake)
        w = self.LOSS_WEIGHTS
        sums, adv, shape = _generator_step_enqueue(self.arrays, self.g, self.d, self.g_opt, True,
                                                   w["content_loss_weighting"], w["topographic_loss_weighting"],
                                                   w["structural_loss_weighting"])
        return out, sums, adv, shape

    def _snapshot(self):
        st = []
        for link, opt in ((self.g, self.g_opt), (self.d, self.d_opt)):
            pers = {k: v.clone() for k, v in getattr(link, "persistent", {}).items()}
            st.append((link.flat.clone(), opt.m.clone(), opt.v.clone(), opt.t, pers))
        return st

    def _restore(self, st):
        for (link, opt), (flat, m, v, t, pers) in zip(((self.g, self.g_opt), (self.d, self.d_opt)), st):
            link.flat.copy_(flat)
            opt.m.copy_(m)
            opt.v.copy_(v)
            opt.t = t
            opt.t_dev.fill_(t)
            for k, val in pers.items():
                link.persistent[k].copy_(val)
            link.mark_updated()
        # The graph re-packs the generator's operand images at its start and the discriminator's after the
        # discriminator update (where the eager step does); the discriminator images its first pass reads are the
        # ones the PREVIOUS step packed, so they must match the restored weights before the first replay:
        # an eval-mode forward re-packs them and changes no state.
        self.d.forward(self.arrays["Y"], train=False)
        torch.cuda.synchronize()

    def accepts(self, input_arrays: Dict[str, object]) -> bool:
        """True when the minibatch has the shapes the graph was captured for."""
        return all(k in input_arrays and tuple(np.shape(input_arrays[k])) == tuple(v.shape)
                   for k, v in self.arrays.items())

    def refresh(self):
        """Call after the weights or Adam state were changed OUTSIDE the graph (eager training steps, load_npz,
        load_state_dict): re-synchronises the device step counts and the discriminator's packed operand images
        (the graph re-packs them only after its own discriminator update)."""
        for opt in (self.g_opt, self.d_opt):
            opt.t_dev.fill_(opt.t)
        self.d.mark_updated()
        self.d.forward(self.arrays["Y"], train=False)

    def step(self, input_arrays: Optional[Dict[str, object]] = None):
        if input_arrays is not None:
            for k, dst in self.arrays.items():
                src = input_arrays[k]
                if src is dst:
                    continue
                if isinstance(src, np.ndarray):
                    src = torch.from_numpy(np.ascontiguousarray(src, dtype=np.float32))
                # host tensors (pinned: asynchronous) go straight into the graph's static input buffers
                dst.copy_(src.array if isinstance(src, Variable) else src, non_blocking=True)
        self.graph.replay()
        for link, opt in ((self.g, self.g_opt), (self.d, self.d_opt)):
            opt.t += 1
            link.mark_updated()      # eager calls made between replays must re-pack their operand images
        torch.cuda.current_stream().synchronize()
        h = self.host.double().numpy()
        gl = _generator_step_finalize(h[2:6], float(h[6]), self.shape, **self.LOSS_WEIGHTS)
        return (float(h[0]), float(h[1])), gl


class ArrayIterator:
    """Minimal SerialIterator stand-in (chainer.iterators.SerialIterator(repeat=True, shuffle=True) as used at
    srgan_train.py:132-166): batches a dict of equally long arrays, counts epochs; batches are always full (the last
    one of an epoch is completed from the next epoch's order, as Chainer does)."""

    def __init__(self, arrays: Dict[str, np.ndarray], batch_size: int, shuffle: bool = True, seed: int = 42):
        self.arrays = arrays
        self.n = len(next(iter(arrays.values())))
        self.batch_size = batch_size
        self.rng = np.random.RandomState(seed)
        self.shuffle = shuffle
        self.epoch = 0
        self._order = self._new_order()
        self._pos = 0

    def _new_order(self):
        return self.rng.permutation(self.n) if self.shuffle else np.arange(self.n)

    def _next_indices(self):
        """chainer SerialIterator(repeat=True): a batch that reaches the end of the epoch is completed from the head
        of the next epoch's (freshly drawn) order, so every batch is full; the epoch counter advances there."""
        idx = self._order[self._pos:self._pos + self.batch_size]
        self._pos += self.batch_size
        if self._pos >= self.n:
            rest = self._pos - self.n
            self.epoch += 1
            self._order = self._new_order()
            if rest > 0:
                idx = np.concatenate([idx, self._order[:rest]])
            self._pos = rest
        return idx

    def next(self):
        idx = self._next_indices()
        return {k: v[idx] for k, v in self.arrays.items()}


class DeviceArrayIterator:
    """On-device minibatch iterator: the whole training set (3826 tiles x 58.4 KB = 235 MB with Y,
    srgan_train.py:132-166) is uploaded once; ``next()`` gathers a shuffled batch with one
    ``dbm_gather_rows_f32`` launch per array, so a step does no host-side array work. Same epoch counting
    and (for the same seed) the same sample order as ``ArrayIterator``."""

    def __init__(self, arrays: Dict[str, object], batch_size: int, shuffle: bool = True, seed: int = 42):
        self.arrays = {k: as_device(v).contiguous() for k, v in arrays.items()}
        lens = {int(v.shape[0]) for v in self.arrays.values()}
        if len(lens) != 1:
            raise ValueError(f"arrays differ in length: {sorted(lens)}")
        self.n = lens.pop()
        if self.n == 0:
            raise ValueError("empty dataset")
        self.batch_size = int(batch_size)
        self.rng = np.random.RandomState(seed)
        self.shuffle = shuffle
        self.epoch = 0
        self._pos = 0
        self._order = self._new_order()

    def _new_order(self):
        return self.rng.permutation(self.n) if self.shuffle else np.arange(self.n)

    _next_indices = ArrayIterator._next_indices

    def next(self):
        idx = torch.from_numpy(np.ascontiguousarray(self._next_indices(), dtype=np.int64)).cuda()   # batch_size int64
        nb = int(idx.numel())
        out = {}
        for k, v in self.arrays.items():
            row = int(v[0].numel())
            dst = ops.empty(nb, *v.shape[1:])
            ops.call("dbm_gather_rows_f32", v.data_ptr(), self.n, idx.data_ptr(), dst.data_ptr(), row, nb, ops.stream())
            out[k] = dst
        return out


def save_model_weights_and_architecture(generator_model, discriminator_model, save_path: str = "model/weights"):
    """srgan_train.py:1333-1383: writes srgan_generator_model_weights.npz and
    srgan_discriminator_model_weights.npz (Chainer key layout, loadable by chainer.serializers.load_npz) and
    srgan_generator_model_architecture.dot. Chainer dumps its autograd graph; there is no autograd graph here, so
    the .dot file lists the generator's layer chain (one node per parametrised layer, in forward order)."""
    import os
    os.makedirs(save_path, exist_ok=True)
    g_path = os.path.join(save_path, "srgan_generator_model_weights.npz")
    d_path = os.path.join(save_path, "srgan_discriminator_model_weights.npz")
    a_path = os.path.join(save_path, "srgan_generator_model_architecture.dot")
    generator_model.save_npz(g_path)
    discriminator_model.save_npz(d_path)
    layers = []
    for k, shp in generator_model._shapes.items():
        if k.endswith("/W"):
            layers.append((k[:-2], tuple(shp)))
    with open(a_path, "w") as fh:
        fh.write("digraph generator {\n  rankdir=TB;\n")
        for i, (name, shp) in enumerate(layers):
            fh.write(f'  n{i} [shape=box, label="{name}\\nW{list(shp)}"];\n')
        stem = [i for i, (nm, _) in enumerate(layers) if nm.startswith("input_block/")]
        body = [i for i in range(len(layers)) if i not in stem]
        for i in stem:
            if body:
                fh.write(f"  n{i} -> n{body[0]};\n")
        for a, b in zip(body[:-1], body[1:]):
            fh.write(f"  n{a} -> n{b};\n")
        fh.write("}\n")
    return g_path, d_path, a_path
