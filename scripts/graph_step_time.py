"""Eager vs CUDA-graph training step (batch 128)."""
import sys, time
import torch
sys.path.insert(0, ".")
from deepbedmap_b200 import train as T

batch = 128
g, g_opt, d, d_opt = T.compile_srgan_model()
for kv in sys.argv[1:]:          # A/B switches of the generator, e.g. resample_from_slab8=0; sm_reserve=N
    k, v = kv.split("=")
    if k == "graph_wgrad_ctas":
        T.GraphedTrainStep.TRUNK_WGRAD_CTAS = int(v)
    elif k == "sm_reserve":
        from deepbedmap_b200 import ops
        ops.call("dbm_set_sm_reserve", int(v))
    else:
        setattr(g, k, type(getattr(g, k))(int(v)))
gen = torch.Generator(device="cuda").manual_seed(42)
r = lambda *s: torch.rand(*s, generator=gen, device="cuda")
arrays = {"X": r(batch, 1, 11, 11), "W1": r(batch, 1, 110, 110), "W2": r(batch, 2, 22, 22), "W3": r(batch, 1, 11, 11),
          "Y": r(batch, 1, 36, 36)}


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def eager():
    T.train_eval_discriminator(arrays, g, d, d_opt, share_generator_forward=True)
    T.train_eval_generator(arrays, g, d, g_opt)


print(f"eager: {timed(eager):.2f} ms per step")
gs = T.GraphedTrainStep(arrays, g, g_opt, d, d_opt)
print(f"graph: {timed(lambda: gs.step(arrays)):.2f} ms per step", gs.step(arrays))
