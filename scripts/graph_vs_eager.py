"""Run-to-run variation of the eager step (fp32 atomics) vs the difference between eager and graphed steps."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from deepbedmap_b200 import train as T

rng = np.random.RandomState(11)
n = 4
batches = [{"X": rng.rand(n, 1, 11, 11), "W1": rng.rand(n, 1, 110, 110), "W2": rng.rand(n, 2, 22, 22),
            "W3": rng.rand(n, 1, 11, 11), "Y": rng.rand(n, 1, 36, 36)} for _ in range(3)]
batches = [{k: torch.as_tensor(v.astype(np.float32)).cuda() for k, v in b.items()} for b in batches]


def eager():
    g, go, d, do = T.compile_srgan_model(num_residual_blocks=1, seed=3)
    out = []
    for b in batches:
        dm = T.train_eval_discriminator(b, g, d, do, share_generator_forward=True)
        out.append((dm, T.train_eval_generator(b, g, d, go)))
    return out, g, d


def graphed():
    g, go, d, do = T.compile_srgan_model(num_residual_blocks=1, seed=3)
    st = T.GraphedTrainStep(batches[0], g, go, d, do)
    return [st.step(b) for b in batches], g, d


a, ga, da = eager()
b, gb, db = eager()
c, gc, dc = graphed()
for i in range(3):
    print("step", i)
    print("  eager A", a[i])
    print("  eager B", b[i])
    print("  graph  ", c[i])
for name, x, y in (("eager A vs eager B", ga, gb), ("eager A vs graph", ga, gc)):
    print(name, "G weights: fraction differing by > 2e-5:", ((x.flat - y.flat).abs() > 2e-5).float().mean().item())
for name, x, y in (("eager A vs eager B", da, db), ("eager A vs graph", da, dc)):
    print(name, "D weights: fraction differing by > 2e-5:", ((x.flat - y.flat).abs() > 2e-5).float().mean().item())
