"""GPU timeline of the training step (bench.py's train leg): wall time per step against the GPU-busy
time (union of kernel intervals, CUPTI through torch.profiler) and the per-kernel totals.
Tells whether the step is bound by the kernels or by the host launching them."""
import collections
import sys
import time

import torch

sys.path.insert(0, ".")
from deepbedmap_b200 import train as T

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
g, g_opt, d, d_opt = T.compile_srgan_model()
gen = torch.Generator(device="cuda").manual_seed(42)
r = lambda *s: torch.rand(*s, generator=gen, device="cuda")
arrays = {"X": r(batch, 1, 11, 11), "W1": r(batch, 1, 110, 110), "W2": r(batch, 2, 22, 22), "W3": r(batch, 1, 11, 11),
          "Y": r(batch, 1, 36, 36)}


def step():
    T.train_eval_discriminator(arrays, g, d, d_opt, share_generator_forward=True)
    T.train_eval_generator(arrays, g, d, g_opt)


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / steps * 1e3
print(f"wall per step (no profiler): {wall:.2f} ms")

from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
iv = sorted((e.time_range.start, e.time_range.end) for e in ev)
busy, cur_s, cur_e = 0.0, None, None
for s, e in iv:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
if cur_e is not None:
    busy += cur_e - cur_s
span = iv[-1][1] - iv[0][0]
print(f"under profiler: {len(ev) / steps:.0f} GPU activities per step, span {span / steps / 1e3:.2f} ms per step, "
      f"GPU busy {busy / steps / 1e3:.2f} ms per step ({100 * busy / span:.0f} %)")
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    k = e.name.replace("void ", "").replace("dbm::", "")[:70]
    agg[k][0] += 1
    agg[k][1] += e.time_range.end - e.time_range.start
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v[1] / steps:9.1f} us/step  n/step={v[0] / steps:6.1f}  avg={v[1] / v[0]:8.1f} us  {k}")
# per-launch durations of the multi-launch tensor-core kernels (one step's worth, in launch order)
per = len(ev) // steps
for key in ("flat_wgrad_kernel", "flat_wgrad_reduce", "flat_conv_kernel", "gemm_f32_kernel", "bn_", "flat_to_nchw", "flat_from_nchw"):
    d = [round(e.time_range.end - e.time_range.start, 1) for e in sorted(ev, key=lambda e: e.time_range.start)[:per] if key in e.name]
    print(f"{key}: {d}")
# the largest gaps in the timeline (host-bound stretches)
gaps = sorted(((iv[i + 1][0] - max(x[1] for x in iv[max(0, i - 3):i + 1]), i) for i in range(len(iv) - 1)), reverse=True)[:10]
print("largest idle gaps (us):", [round(gp, 1) for gp, _ in gaps])
