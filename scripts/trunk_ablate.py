"""Times the persistent trunk kernel alone on a batch of interior continent tiles (286x286 trunk
pixels each) for the paired / unpaired plans and the kernel's ablation masks (dbm_debug_set(3, mask):
1 no dependency wait, 2 no epilogue memory traffic, 4 no TMA loads -- results are invalid under a
mask, only the time is meaningful). Prints TFLOP/s of algorithmic work."""
import sys
import torch
sys.path.insert(0, ".")
from deepbedmap_b200 import GeneratorModel, _lib, ops
from bench import ClockSampler

ns = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [4]
masks = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 2, 4, 7]
plans = (True, False) if len(sys.argv) <= 3 else (True,)
H = W = 286
m = GeneratorModel(precision="bf16")
for n, paired in [(n, pl) for n in ns for pl in plans]:
    m.paired_trunk = paired
    ws = m._trunk_workspace(n, H, W)
    ws["s0"].normal_()
    for mask in masks:
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
        print(f"n={n} paired={paired} mask={mask}: {ms:7.3f} ms  {ws['flops'] / ms / 1e9:7.1f} TFLOP/s  "
              f"sm_mhz={c.get('sm_mhz')} power={c.get('power_w_max')} {c.get('reasons')}", flush=True)
    _lib.call("dbm_debug_set", 3, 0)
    if "prof" in sys.argv:
        # per-pass cycle counters (summed over CTAs): where the MMA issuer and the epilogue wait
        nl = len(ws["layers"])
        prof = torch.zeros(nl * 8, dtype=torch.int64, device="cuda")
        _lib.call("dbm_debug_set_ptr", 1, prof.data_ptr())
        _lib.call("dbm_debug_set", 3, masks[0])
        m._run_trunk(ws, n, H, W)
        torch.cuda.synchronize()
        _lib.call("dbm_debug_set", 3, 0)
        _lib.call("dbm_debug_set_ptr", 1, None)
        pr = prof.view(nl, 8).cpu().double()
        kinds = {}
        for L, ly in enumerate(ws["layers"]):
            key = (ly[7], ly[8], ly[15])  # cin, cout, cout_main
            kinds.setdefault(key, []).append(L)
        print(f"n={n} paired={paired} mask={masks[0]}: per pass kind, cycles per item: MMA[wait tmem, wait operands, total] "
              f"EPI[wait acc, read-out, stores+fence]")
        for key, Ls in kinds.items():
            a = pr[Ls].sum(0)
            cnt = a[3].item()
            mm = (9 * 4 * key[0] / 16) * (46.4 if key[1] == 32 else 53.9)
            print(f"  cin={key[0]:3d} cout={key[1]} main={key[2]}: passes={len(Ls):3d} items={int(cnt)}  "
                  f"MMA {a[0].item()/cnt:8.0f} {a[1].item()/cnt:8.0f} {a[2].item()/cnt:8.0f} (ideal {mm:6.0f})  "
                  f"EPI {a[4].item()/cnt:8.0f} {a[5].item()/cnt:8.0f} {a[6].item()/cnt:8.0f}")
