"""Training step (D-step + G-step, batch 128) for both dense-block widths of the reference's search space
(inter_channels 32 / 64, srgan_train.py:283-284) and both training precisions: eager ms per step, and the
CUDA-graph replay for the tensor-core path."""
import sys, time
import torch
sys.path.insert(0, ".")
from deepbedmap_b200 import train as T

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 128
gen = torch.Generator(device="cuda").manual_seed(42)
r = lambda *s: torch.rand(*s, generator=gen, device="cuda")
arrays = {"X": r(batch, 1, 11, 11), "W1": r(batch, 1, 110, 110), "W2": r(batch, 2, 22, 22), "W3": r(batch, 1, 11, 11),
          "Y": r(batch, 1, 36, 36)}


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


for ic in (32, 64):
    for prec in ("bf16", "fp32"):
        g, g_opt, d, d_opt = T.compile_srgan_model(inter_channels=ic, train_precision=prec)

        def eager():
            T.train_eval_discriminator(arrays, g, d, d_opt, share_generator_forward=True)
            return T.train_eval_generator(arrays, g, d, g_opt)

        line = f"inter_channels={ic} train_precision={prec}: eager {timed(eager):7.2f} ms per step"
        if prec == "bf16":
            gs = T.GraphedTrainStep(arrays, g, g_opt, d, d_opt)
            line += f", graph replay {timed(lambda: gs.step(arrays)):6.2f} ms; metrics {gs.step(arrays)}"
        print(line, flush=True)
        del g, g_opt, d, d_opt
        torch.cuda.empty_cache()
