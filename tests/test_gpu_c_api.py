"""Model-level C entry points (include/deepbedmap_b200.h, dbm_gen_*): a forward run by a C program that links the
library and never touches the Python shim, against the fp64 oracle and -- bit for bit -- against the Python class."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import deepbedmap_oracle as O  # noqa: E402  (test infrastructure)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


@pytest.mark.parametrize("nb,n,h,w", [(2, 2, 21, 38), (1, 1, 40, 37)])
def test_c_host_runs_the_generator_forward(tmp_path, nb, n, h, w):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "gen_forward_main")
    lib_dir = os.path.join(ROOT, "deepbedmap_b200")
    subprocess.run([nvcc, "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "c_api", "gen_forward_main.cu"), "-L", lib_dir, "-ldeepbedmap_b200",
                    "-Xlinker", "-rpath", "-Xlinker", lib_dir], check=True)
    params = O.init_generator_params(nb, seed=0, bias_std=0.1, scale=0.7)
    ins = O.synthetic_inputs(n, h, w)
    np.concatenate([v.reshape(-1) for v in params.values()]).astype(np.float32).tofile(tmp_path / "params.bin")
    np.concatenate([a.reshape(-1) for a in ins]).astype(np.float32).tofile(tmp_path / "inputs.bin")
    out = subprocess.run([exe, str(nb), "0.1", str(n), str(h), str(w), str(tmp_path / "params.bin"),
                          str(tmp_path / "inputs.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    print(out.stdout.strip())
    y = np.fromfile(tmp_path / "out.bin", np.float32).reshape(n, 1, 4 * (h - 2), 4 * (w - 2))
    ref = O.generator_forward_numpy(params, *ins, num_residual_blocks=nb)
    assert rel_l2(y, ref) < 2e-2
    # the Python class drives the same entry point (and, with c_model_api off, composes the same kernels itself)
    from deepbedmap_b200 import GeneratorModel
    m = GeneratorModel(num_residual_blocks=nb, precision="bf16")
    for k, v in params.items():
        m.set_param(k, v)
    m.local_trunk = False
    assert m._c_forward_applies()
    via_class = m.forward(*ins).numpy()
    assert np.array_equal(via_class, y)
    m.c_model_api = False
    composed = m.forward(*ins).numpy()
    assert np.array_equal(composed, y)


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
def test_c_api_follows_weight_updates(precision):
    """The handle views the class's flat parameter buffer: a set_param / optimizer update must reach the packed operands
    (bf16 and split-bf16 images alike)."""
    from deepbedmap_b200 import GeneratorModel
    params = O.init_generator_params(1, seed=0, bias_std=0.1, scale=0.7)
    m = GeneratorModel(num_residual_blocks=1, precision=precision)
    for k, v in params.items():
        m.set_param(k, v)
    m.local_trunk = False
    ins = O.synthetic_inputs(1, 30, 20)
    y0 = m.forward(*ins).array.clone()
    key = "residual_network/0/residual_dense_block2/conv_layer3/W"
    m.set_param(key, params[key] * 1.5)
    y1 = m.forward(*ins).array.clone()
    assert not torch.equal(y0, y1)
    m.c_model_api = False
    assert torch.equal(m.forward(*ins).array, y1)


def test_c_model_api_wide_dense_blocks_equal_python_composition():
    """inter_channels = 64 (srgan_train.py:283-284) through dbm_gen_forward (the handle builds the unpaired pass table
    and the 320-channel dense-block workspace) against the same kernels composed by the Python class, bit for bit, and
    against the oracle."""
    from deepbedmap_b200 import GeneratorModel
    nb = 2
    params = O.init_generator_params(nb, seed=0, bias_std=0.1, scale=0.5, inter_channels=64)
    m = GeneratorModel(num_residual_blocks=nb, precision="bf16", inter_channels=64)
    for k, v in params.items():
        m.set_param(k, v)
    ins = O.synthetic_inputs(2, 37, 44)
    assert m._c_forward_applies()
    y = m.forward(*ins).numpy()
    m.c_model_api = False
    assert np.array_equal(m.forward(*ins).numpy(), y)
    ref = O.generator_forward_numpy(params, *ins, num_residual_blocks=nb)
    assert rel_l2(y, ref) < 2e-2


def test_c_host_runs_the_split_bf16_forward(tmp_path):
    """dbm_gen_set_precision(gen, 1): the split-bf16 path (fp32 stem, split trunk + upsample convs, fp32 deformable
    layers) composed in C++: fp32-grade against the fp64 oracle from a C-only host, and bit-identical to the Python
    class, whether it drives the handle or composes the per-op entry points itself."""
    nb, n, h, w = 2, 2, 21, 38
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "gen_forward_main")
    lib_dir = os.path.join(ROOT, "deepbedmap_b200")
    subprocess.run([nvcc, "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "c_api", "gen_forward_main.cu"), "-L", lib_dir, "-ldeepbedmap_b200",
                    "-Xlinker", "-rpath", "-Xlinker", lib_dir], check=True)
    params = O.init_generator_params(nb, seed=0, bias_std=0.1, scale=0.7)
    ins = O.synthetic_inputs(n, h, w)
    np.concatenate([v.reshape(-1) for v in params.values()]).astype(np.float32).tofile(tmp_path / "params.bin")
    np.concatenate([a.reshape(-1) for a in ins]).astype(np.float32).tofile(tmp_path / "inputs.bin")
    out = subprocess.run([exe, str(nb), "0.1", str(n), str(h), str(w), str(tmp_path / "params.bin"),
                          str(tmp_path / "inputs.bin"), str(tmp_path / "out.bin"), "1"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    y = np.fromfile(tmp_path / "out.bin", np.float32).reshape(n, 1, 4 * (h - 2), 4 * (w - 2))
    ref = O.generator_forward_numpy(params, *ins, num_residual_blocks=nb)
    print(f"C host, bf16x3: rel_l2 vs fp64 oracle {rel_l2(y, ref):.2e}")
    assert rel_l2(y, ref) < 5e-5
    from deepbedmap_b200 import GeneratorModel
    m = GeneratorModel(num_residual_blocks=nb, precision="bf16x3")
    for k, v in params.items():
        m.set_param(k, v)
    assert np.array_equal(m.forward(*ins).numpy(), y)           # through the handle
    m.c_model_api = False
    assert np.array_equal(m.forward(*ins).numpy(), y)           # composed from per-op calls in model.py
