"""Localises trunk-kernel discrepancies: per-layer launches vs persistent (unpaired / paired)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from deepbedmap_b200 import GeneratorModel, _lib, ops
from oracle import deepbedmap_oracle as O

nb, n, h, w = 1, 2, 11, 11
if len(sys.argv) > 4:
    nb, n, h, w = map(int, sys.argv[1:5])
H, W = h - 2, w - 2
m = GeneratorModel(num_residual_blocks=nb, precision="bf16", init_scale=0.7)
torch.manual_seed(0)
s0 = torch.randn(n, 16, H, W, 8, device="cuda").bfloat16()

def run(persistent, paired, dbg=0):
    m.persistent_trunk, m.paired_trunk = persistent, paired
    ws = m._trunk_workspace(n, H, W)
    ws["s0"].copy_(s0)
    for t in ws["cat"] + ws["f32"] + [ws["a1_f32"], ws["u1"]]:
        t.zero_()
    _lib.call("dbm_debug_set", 3, dbg)
    m._run_trunk(ws, n, H, W)
    _lib.call("dbm_debug_set", 3, 0)
    torch.cuda.synchronize()
    return {"cat0": ws["cat"][0].float().clone(), "cat1": ws["cat"][1].float().clone(), "u1": ws["u1"].float().clone(),
            "a1": ws["a1_f32"].clone()}

ref = run(False, False)
for name, args in (("persistent-unpaired", (True, False)), ("persistent-paired", (True, True)),
                   ("persistent-paired-allfence", (True, True, 8)), ("persistent-unpaired-allfence", (True, False, 8))):
    got = run(*args)
    for k in ref:
        d = (got[k] - ref[k]).norm() / (ref[k].norm() + 1e-30)
        print(f"{name:30s} {k:5s} rel diff {float(d):.3e}  ref norm {float(ref[k].norm()):.3e}")
    if nb == 1:
        c = got["cat0"]; r = ref["cat0"]
        for slab0 in range(0, c.shape[1], 4):
            d = (c[:, slab0:slab0 + 4] - r[:, slab0:slab0 + 4]).norm() / (r[:, slab0:slab0 + 4].norm() + 1e-30)
            print(f"    cat0 slabs {slab0}-{slab0+3}: {float(d):.3e}")
