"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads and exports every
symbol include/deepbedmap_b200.h declares; host-side inventories agree with the oracle."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header="deepbedmap_b200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dbm_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from deepbedmap_b200 import _lib, build
    build.build()
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    # every declared entry point has a ctypes signature (and vice versa)
    assert set(names) - {"dbm_last_error"} == set(_lib.SIGNATURES)
    assert lib.dbm_version() >= 100
    # the boundary header declares no tuning / ablation hook; those live in their own header (and the microbenchmark in
    # its own library)
    assert not any("debug" in n for n in names)
    tuning = set(_declared_symbols("deepbedmap_b200_tuning.h"))
    assert tuning - {"dbm_debug_umma_rate"} == set(_lib.TUNING_SIGNATURES)
    assert not hasattr(lib, "dbm_debug_umma_rate")


def test_signature_arity_matches_header():
    from deepbedmap_b200 import _lib
    text = open(os.path.join(ROOT, "include", "deepbedmap_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, args in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\((.*?)\)\s*;" % name, text, flags=re.S)
        assert m, name
        decl = m.group(1).strip()
        n = 0 if decl in ("", "void") else decl.count(",") + 1
        assert n == len(args), f"{name}: header has {n} args, ctypes table {len(args)}"


def test_layout_matches_oracle_inventory():
    from deepbedmap_b200 import layout
    from oracle import deepbedmap_oracle as O
    for nb in (1, 12):
        assert list(layout.generator_shapes(nb).items()) == list(O.generator_param_shapes(nb).items())
    assert list(layout.discriminator_shapes().items()) == list(O.discriminator_param_shapes().items())
    assert sum(int(np.prod(s)) for s in layout.generator_shapes().values()) == 8907749
    assert sum(int(np.prod(s)) for s in layout.discriminator_shapes().values()) == 10370761
    assert layout.infer_num_residual_blocks(layout.generator_shapes(7)) == 7


def test_tile_plan_matches_oracle():
    from deepbedmap_b200.tiler import tile_plan
    from oracle import deepbedmap_oracle as O
    a = tile_plan()
    b = O.tile_plan()
    assert len(a) == len(b) == 396
    for t, r in zip(a, b):
        assert t[:4] == r[:4] and (t[4], t[5]) == (r[4].start, r[4].stop) and (t[6], t[7]) == (r[5].start, r[5].stop)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "deepbedmap_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src, f"{f} references the oracle"


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from deepbedmap_b200 import GeneratorModel
    with pytest.raises(RuntimeError):
        GeneratorModel()


def test_flat_layout_geometry_host_formula_matches_library():
    """The flat-padded layout of the tensor-core training trunk: host tables (flat.py) and kernels
    (umma_flat.cu) must agree on {P, tiles, G0, Pg, R}; dbm_flat_geometry is host-only arithmetic."""
    from deepbedmap_b200 import flat
    for n, h, w in [(1, 1, 1), (128, 9, 9), (3, 9, 9), (5, 7, 12), (1, 20, 33), (2, 150, 170)]:
        g = flat.geometry(n, h, w)
        assert g == flat.geometry_host(n, h, w)
        assert g["G0"] >= w + 3 and g["G0"] % 8 == 0 and g["Pg"] == 2 * g["G0"] + 128 * g["tiles"]
        assert 128 * g["tiles"] >= g["P"] == n * (h + 2) * (w + 2)
    assert flat.split_blocks(121, 3) == [(0, 41), (41, 40), (81, 40)]
    assert sum(k for _, k in flat.split_blocks(7, 16)) == 7
    assert flat.chunk_channels(160) == [(0, 128), (128, 32)]
    # d(x_{j+1}) -> conv5 output gradient scale: beta, beta, beta^2 within every RRDB (srgan_train.py:358, 402)
    assert [round(d["g5_scale"], 6) for d in flat.rdb_plan(6, 0.1)] == [0.1, 0.1, 0.01, 0.1, 0.1, 0.01]


def test_table_records_match_the_kernels_struct_sizes():
    """The device tables the host builds (numpy record dtypes) must have the byte size the kernels static_assert
    for their structs -- a drifted field would shift every pointer behind it."""
    from deepbedmap_b200 import flat, model
    csrc = os.path.join(ROOT, "deepbedmap_b200", "csrc")
    sizes = {}
    for f in os.listdir(csrc):
        for name, size in re.findall(r"static_assert\(sizeof\((\w+)\) == (\d+)", open(os.path.join(csrc, f)).read()):
            sizes[name] = int(size)
    pairs = {"PackEntry": model.PACK_ENTRY_DTYPE, "TrunkLayer": model.TRUNK_LAYER_DTYPE, "FlatEpiBlock": flat.EPI_DTYPE,
             "FlatLaunch": flat.LAUNCH_DTYPE, "WgradUnit": flat.WGRAD_UNIT_DTYPE, "WgradReduce": flat.WGRAD_REDUCE_DTYPE,
             "LocalPass": flat.LOCAL_PASS_DTYPE}
    assert set(pairs) <= set(sizes), sorted(set(pairs) - set(sizes))
    for name, dt in pairs.items():
        assert dt.itemsize == sizes[name], (name, dt.itemsize, sizes[name])


def test_image_resident_trunk_applicability():
    """flat.local_trunk_fits: a padded image must be one M=128 MMA tile (the reference's 11x11 windows are)."""
    from deepbedmap_b200 import flat
    assert flat.local_trunk_fits(9, 9) and flat.local_trunk_fits(6, 14) and flat.local_trunk_fits(1, 1)
    assert not flat.local_trunk_fits(10, 10) and not flat.local_trunk_fits(286, 286)
    g = flat.geometry_host(128, 9, 9)
    assert g["P"] == 128 * 121 and g["G0"] >= 12 and g["Pg"] == 2 * g["G0"] + 128 * g["tiles"]
