"""Parameter inventory of the generator / discriminator in the reference's Chainer ``.npz`` key
layout (chainer.serializers.save_npz of the links defined at srgan_train.py:201-699; call sites
srgan_train.py:1355-1361, deepbedmap.py:408)."""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

# (out_channels, ksize, stride) of DiscriminatorModel.conv_layer0..9 (srgan_train.py:617-634)
DISC_CONVS = [(64, 3, 1), (64, 4, 2), (128, 3, 1), (128, 4, 2), (128, 3, 1),
              (256, 4, 2), (256, 3, 1), (512, 4, 2), (512, 3, 1), (512, 4, 2)]


def generator_shapes(num_residual_blocks: int = 12, inter_channels: int = 32, out_channels: int = 1):
    g = inter_channels
    s = OrderedDict()
    for name, shp in (("X", (32, 1, 3, 3)), ("W1", (32, 1, 30, 30)), ("W2", (32, 2, 6, 6)), ("W3", (32, 1, 3, 3))):
        s[f"input_block/conv_on_{name}/W"] = shp
        s[f"input_block/conv_on_{name}/b"] = (32,)
    s["pre_residual_conv_layer/W"] = (64, 128, 3, 3)
    s["pre_residual_conv_layer/b"] = (64,)
    for i in range(num_residual_blocks):
        for r in (1, 2, 3):
            p = f"residual_network/{i}/residual_dense_block{r}"
            for k in (1, 2, 3, 4):
                s[f"{p}/conv_layer{k}/W"] = (g, 64 + (k - 1) * g, 3, 3)
                s[f"{p}/conv_layer{k}/b"] = (g,)
            s[f"{p}/conv_layer5/W"] = (64, 64 + 4 * g, 3, 3)
            s[f"{p}/conv_layer5/b"] = (64,)
    for name in ("post_residual_conv_layer", "post_upsample_conv_layer_1", "post_upsample_conv_layer_2"):
        s[f"{name}/W"] = (64, 64, 3, 3)
        s[f"{name}/b"] = (64,)
    for name, oc in (("final_conv_layer1", 64), ("final_conv_layer2", out_channels)):
        s[f"{name}/offset_conv/W"] = (18, 64, 3, 3)
        s[f"{name}/offset_conv/b"] = (18,)
        s[f"{name}/deform_conv/W"] = (oc, 64, 3, 3)
        s[f"{name}/deform_conv/b"] = (oc,)
    return s


def discriminator_shapes():
    s = OrderedDict()
    cin = 1
    for i, (cout, k, _) in enumerate(DISC_CONVS):
        s[f"conv_layer{i}/W"] = (cout, cin, k, k)
        if i == 0:
            s["conv_layer0/b"] = (cout,)
        else:
            s[f"batch_norm{i}/gamma"] = (cout,)
            s[f"batch_norm{i}/beta"] = (cout,)
        cin = cout
    s["linear_1/W"] = (100, 512)
    s["linear_1/b"] = (100,)
    s["linear_2/W"] = (1, 100)
    s["linear_2/b"] = (1,)
    return s


def discriminator_persistents():
    """avg_mean / avg_var / N of every L.BatchNormalization (serialised with the params)."""
    s = OrderedDict()
    for i, (cout, _, _) in enumerate(DISC_CONVS):
        if i:
            s[f"batch_norm{i}/avg_mean"] = (cout,)
            s[f"batch_norm{i}/avg_var"] = (cout,)
    return s


def he_normal(rng: np.random.RandomState, shape, scale: float = 0.1) -> np.ndarray:
    """chainer.initializers.HeNormal(scale=0.1, fan_option="fan_in") (srgan_train.py:220)."""
    fan_in = int(np.prod(shape[1:]))
    return rng.normal(0.0, scale * math.sqrt(2.0 / fan_in), size=shape).astype(np.float32)


def init_values(shapes, seed: int, scale: float = 0.1):
    rng = np.random.RandomState(seed)
    out = OrderedDict()
    for k, shp in shapes.items():
        if k.endswith("/W"):
            out[k] = he_normal(rng, shp, scale)
        elif k.endswith("/gamma"):
            out[k] = np.ones(shp, np.float32)
        else:
            out[k] = np.zeros(shp, np.float32)
    return out


def infer_num_residual_blocks(keys) -> int:
    """The .npz does not store the block count (the reference gets it from Comet,
    deepbedmap.py:397-405); infer it from the highest residual_network/<i> index."""
    idx = [int(k.split("/")[1]) for k in keys if k.startswith("residual_network/")]
    return max(idx) + 1 if idx else 0
