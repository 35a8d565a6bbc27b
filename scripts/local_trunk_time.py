"""Times the training trunk forward + data-gradient chain (batch of 9x9 tiles): image-resident kernels
(1 or 2 images per CTA) vs the flat chain."""
import sys
import torch
sys.path.insert(0, ".")
from deepbedmap_b200 import GeneratorModel
from deepbedmap_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
m = GeneratorModel(num_residual_blocks=12)
ft = m._flat_trunk(n, 9, 9)
a0 = torch.randn(n, 128, 9, 9, device="cuda")
st = ops.stream()


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10


ft.forward(a0)
m._pack(m.PACK_TRAIN_CHAIN)
for label, local, group in (("flat chain", False, 0), ("image-resident, 1 image/CTA", True, 1), ("image-resident, 2 images/CTA in lock step", True, 2),
                            ("image-resident, SOLO (2 CTAs/SM)", True, 3)):
    ft.local = local
    ops.call("dbm_local_debug_set", group)
    if local:
        f = lambda: ops.call("dbm_trunk_local_fwd", ft.local_dev.data_ptr(), ft.n_local, n, 9, 9, ft.s0.data_ptr(),
                             ft.x_scratch[0].data_ptr(), ft.x_scratch[1].data_ptr(), st)
        b = lambda: ops.call("dbm_trunk_local_bwd", ft.local_bwd_dev.data_ptr(), ft.n_local_bwd, n, 9, 9,
                             ft.gpost.data_ptr(), ft.x_scratch[1].data_ptr(), st)
    else:
        f = lambda: ft._chain(ft.fwd, ft.fwd_dev)
        b = lambda: ft._chain(ft.bwd, ft.bwd_dev)
    tf, tb = timed(f), timed(b)
    print(f"n={n} {label}: forward {tf:.3f} ms ({ft.flops_fwd / tf / 1e9:.0f} TFLOP/s), data gradient {tb:.3f} ms "
          f"({ft.flops_fwd / tb / 1e9:.0f} TFLOP/s)")
ops.call("dbm_local_debug_set", 0)
