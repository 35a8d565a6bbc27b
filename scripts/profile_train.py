"""Runs ESRGAN training steps (D-step + G-step, batch 128) for profiling under ncu."""
import sys
import torch
sys.path.insert(0, ".")
from deepbedmap_b200 import train as T

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
g, g_opt, d, d_opt = T.compile_srgan_model()
gen = torch.Generator(device="cuda").manual_seed(42)
r = lambda *s: torch.rand(*s, generator=gen, device="cuda")
arrays = {"X": r(batch, 1, 11, 11), "W1": r(batch, 1, 110, 110), "W2": r(batch, 2, 22, 22), "W3": r(batch, 1, 11, 11),
          "Y": r(batch, 1, 36, 36)}
import time
for i in range(steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dl = T.train_eval_discriminator(arrays, g, d, d_opt, share_generator_forward=True)   # as trainer() / bench.py
    torch.cuda.synchronize(); t1 = time.perf_counter()
    gl = T.train_eval_generator(arrays, g, d, g_opt)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"step {i}: D-step {1e3 * (t1 - t0):.1f} ms, G-step {1e3 * (t2 - t1):.1f} ms", flush=True)
print("ok", dl, gl)
