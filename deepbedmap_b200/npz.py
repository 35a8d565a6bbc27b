"""Reader / writer for the reference's Chainer ``.npz`` weight files
(chainer.serializers.save_npz / load_npz; srgan_train.py:1355-1361, 1566-1574; deepbedmap.py:408).
Keys are '/'-joined link paths, values float32 arrays; BatchNormalization persistents
(avg_mean, avg_var, N) are stored next to the parameters."""
from __future__ import annotations

import numpy as np
import torch


def load_npz(file, model, strict: bool = True) -> None:
    with np.load(file) as f:
        keys = set(f.files)
        missing = [k for k in model.p if k not in keys]
        if strict and missing:
            raise KeyError(f"{file}: missing parameters {missing[:4]}{'...' if len(missing) > 4 else ''}")
        for k in model.p:
            if k in keys:
                model.set_param(k, f[k])
        persistent = getattr(model, "persistent", None)
        if persistent is not None:
            for k, t in persistent.items():
                if k in keys:
                    t.copy_(torch.from_numpy(np.asarray(f[k], np.float32)))
                elif strict:
                    raise KeyError(f"{file}: missing persistent {k}")
            for i in getattr(model, "bn_N", {}):
                k = f"batch_norm{i}/N"
                if k in keys:
                    model.bn_N[i] = int(f[k])
        unexpected = [k for k in keys if k not in model.p and not (persistent is not None and (
            k in persistent or k.endswith("/N")))]
        if strict and unexpected:
            raise KeyError(f"{file}: unexpected keys {unexpected[:4]} (wrong num_residual_blocks?)")


def save_npz(file, model, compression: bool = True) -> None:
    out = {k: v for k, v in model.state_dict().items()}
    persistent = getattr(model, "persistent", None)
    if persistent is not None:
        for k, t in persistent.items():
            out[k] = t.detach().cpu().numpy()
        for i, n in getattr(model, "bn_N", {}).items():
            out[f"batch_norm{i}/N"] = np.asarray(n, dtype=np.int64)
    (np.savez_compressed if compression else np.savez)(file, **out)


def peek_num_residual_blocks(file) -> int:
    from .layout import infer_num_residual_blocks
    with np.load(file) as f:
        return infer_num_residual_blocks(f.files)
