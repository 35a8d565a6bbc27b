"""Per-stream GPU timeline of the GRAPH-REPLAYED training step (GraphedTrainStep, the form bench.py times): CUPTI kernel
records through torch.profiler, grouped by CUDA stream. Prints, per step: the span, each stream's busy time and its
kernels' totals -- the stream whose busy time is closest to the span carries the critical path -- and the idle time
of the main stream split by which stream was running meanwhile."""
import collections
import sys

import torch

sys.path.insert(0, ".")
from deepbedmap_b200 import train as T

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
g, g_opt, d, d_opt = T.compile_srgan_model()
gen = torch.Generator(device="cuda").manual_seed(42)
r = lambda *s: torch.rand(*s, generator=gen, device="cuda")
arrays = {"X": r(batch, 1, 11, 11), "W1": r(batch, 1, 110, 110), "W2": r(batch, 2, 22, 22), "W3": r(batch, 1, 11, 11),
          "Y": r(batch, 1, 36, 36)}
gs = T.GraphedTrainStep(arrays, g, g_opt, d, d_opt)
for _ in range(5):
    gs.step(arrays)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    gs.step(arrays)
e1.record()
torch.cuda.synchronize()
print(f"graph replay: {e0.elapsed_time(e1) / 20:.3f} ms per step")

from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        gs.step(arrays)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower()]
ev.sort(key=lambda e: e.time_range.start)
span = (ev[-1].time_range.end - ev[0].time_range.start) / steps
print(f"under profiler: {len(ev) / steps:.0f} kernels per step, span {span / 1e3:.3f} ms per step")
streams = collections.defaultdict(list)
for e in ev:
    streams[getattr(e, "device_resource_id", getattr(e, "stream", 0))].append(e)
for sid, es in sorted(streams.items(), key=lambda kv: -sum(e.time_range.end - e.time_range.start for e in kv[1])):
    busy = sum(e.time_range.end - e.time_range.start for e in es) / steps
    print(f"\nstream {sid}: {len(es) / steps:.0f} kernels per step, busy {busy / 1e3:.3f} ms per step")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in es:
        k = e.name.replace("void ", "").replace("dbm::", "")[:60]
        agg[k][0] += 1
        agg[k][1] += e.time_range.end - e.time_range.start
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"   {v[1] / steps:8.1f} us/step  n={v[0] / steps:5.1f}  avg={v[1] / v[0]:7.1f} us  {k}")
# one step in time order: (start offset, duration, stream, name) for kernels longer than 40 us
per = len(ev) // steps
one = ev[per * (steps - 1):]
t0 = one[0].time_range.start
print("\nlast step, kernels >= 40 us in start order (offset us, duration us, stream, name):")
for e in one:
    dur = e.time_range.end - e.time_range.start
    if dur >= 40:
        print(f"   {e.time_range.start - t0:8.1f} {dur:7.1f}  s{getattr(e, 'device_resource_id', 0)}  "
              f"{e.name.replace('void ', '').replace('dbm::', '')[:60]}")
print(f"   step ends at {one[-1].time_range.end - t0:.1f} us")
if len(sys.argv) > 3:   # full kernel list of the last step (every launch): offset, duration, stream, name
    with open(sys.argv[3], "w") as fh:
        for e in one:
            fh.write(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f} "
                     f"s{getattr(e, 'device_resource_id', 0)} {e.name.replace('void ', '').replace('dbm::', '')[:70]}\n")
