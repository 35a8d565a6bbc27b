"""Times the persistent trunk kernel alone on a batch of interior continent tiles (286x286 trunk
pixels each) for the paired / unpaired plans and the kernel's ablation masks (dbm_debug_set(3, mask):
1 no dependency wait, 2 no epilogue memory traffic, 4 no TMA loads -- results are invalid under a
mask, only the time is meaningful). Prints TFLOP/s of algorithmic work."""
import sys
import torch
sys.path.insert(0, ".")
from deepbedmap_b200 import GeneratorModel, _lib, ops
from bench import ClockSampler

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H = W = 286
m = GeneratorModel(precision="bf16")
for paired in (True, False):
    m.paired_trunk = paired
    ws = m._trunk_workspace(n, H, W)
    ws["s0"].normal_()
    for mask in (0, 2, 4, 7):
        _lib.call("dbm_debug_set", 3, mask)
        for _ in range(2):
            m._run_trunk(ws, n, H, W)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 60
        with ClockSampler(0) as cs:
            e0.record()
            for _ in range(reps):
                m._run_trunk(ws, n, H, W)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        c = cs.summary()
        print(f"paired={paired} mask={mask}: {ms:7.3f} ms  {ws['flops'] / ms / 1e9:7.1f} TFLOP/s  "
              f"sm_mhz={c.get('sm_mhz')} power={c.get('power_w_max')} {c.get('reasons')}", flush=True)
    _lib.call("dbm_debug_set", 3, 0)
