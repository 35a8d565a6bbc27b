#!/usr/bin/env python
"""Benchmark of the DeepBedMap ESRGAN hot path on B200 (contract: see DESIGN.md "Measurement").

Headline workload (BASELINE.json configs[2]): whole-Antarctic tiled inference on a synthetic
continent-sized grid (X 1x1x4502x5502, W1 1x1x45020x55020, W2 1x2x9004x11004, W3 1x1x4502x5502 ->
18000 x 22000 px at 250 m = 396 Mpx), reference tile geometry, 12 RRDB generator with
random-init weights. One "step" = one whole-continent pass. With N GPUs the 396 tiles are split
into N contiguous runs (strong scaling, one final gather).

  value : Mpx/s with the grids already resident in HBM, predictions left on the device
  e2e   : Mpx/s through predict_continent() with pinned HOST grids in and a HOST DEM out
  roofline : tcgen05 3x3-conv kernel, algorithmic FLOPs / CUDA-event time vs measured bf16 peak
  cpu_baseline / --impl reference : the reference graph restated in torch-CPU fp32 (Chainer is
      not installable in this image), timed on this box's host cores on a bounded sample
  train : secondary metric, ESRGAN train steps/s (D-step + G-step, batch 128 per GPU)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MPX = 396.0  # 18000 x 22000 output pixels
TRUNK_DRAM_BYTES_PER_FLOP = 24.515e9 / 5.717e12  # measured, see instrumented_roofline()
FULL = dict(final_shape=(18000, 22000), ary_shape=(1000, 1000), grid=(4502, 5502))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1383.0), d.get("hbm_gbs", 6547.2), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [s.strip() for s in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]),
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows), "reasons": reasons}


def synth_grids_device(grid, seed=42):
    """Synthetic continent in the physical regime of SURVEY §8(d) config 3; identical on every rank."""
    H, W = grid
    g = torch.Generator(device="cuda").manual_seed(seed)
    X = (torch.randn(1, 1, H, W, generator=g, device="cuda") * 800.0 - 500.0).clamp_(-5000.0, 4500.0)
    W1 = torch.rand(1, 1, 10 * H, 10 * W, generator=g, device="cuda") * 4000.0
    W2 = torch.randn(1, 2, 2 * H, 2 * W, generator=g, device="cuda") * 200.0
    W3 = torch.rand(1, 1, H, W, generator=g, device="cuda") * 1000.0
    return X, W1, W2, W3


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        import datetime
        # short watchdog: a mismatched collective must fail in minutes, not burn the GPU box for ten
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms, world):
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])
    return ms


# --------------------------------------------------------------------------------------------
# CPU baseline (the oracle port; the only place besides tests where oracle/ is executed)
# --------------------------------------------------------------------------------------------
def cpu_reference_rate(crop=192, reps=3, warm=1, nb=12):
    """Times the reference graph (torch-CPU fp32 restatement) on one crop x crop lowres window
    and extrapolates to the continent by computed-pixel count. Returns (Mpx/s, cores, sample)."""
    from oracle import deepbedmap_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = O.to_torch(O.init_generator_params(nb, seed=0), torch.float32)
    ins = [torch.as_tensor(a) for a in O.synthetic_inputs(1, crop, crop, regime="physical")]
    with torch.no_grad():
        for _ in range(warm):
            O.generator_forward(params, *ins, num_residual_blocks=nb)
        t0 = time.perf_counter()
        for _ in range(reps):
            O.generator_forward(params, *ins, num_residual_blocks=nb)
        dt = (time.perf_counter() - t0) / reps
    px_computed = (4 * (crop - 2)) ** 2
    # continent: sum over the 396 tiles of the pixels the generator computes (incl. the halo)
    total = sum(16 * (y1 - y0 - 2) * (x1 - x0 - 2) for (y0, y1, x0, x1, _, _) in O.tile_plan())
    t_cont = dt * total / px_computed
    sample = (f"{reps} forward(s) of one {crop}x{crop} lowres crop (-> {4 * (crop - 2)}^2 px), {dt:.2f} s each, "
              f"extrapolated x{total / px_computed:.0f} by computed-pixel count to the 396-tile continent")
    return MPX / t_cont, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    sample = ""
    cores = os.cpu_count()
    for i in range(args.warmup + args.steps):
        v, cores, sample = cpu_reference_rate(crop=args.cpu_crop, reps=1, warm=0)
        if i >= args.warmup:
            vals.append(v)
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": "continent inference output Mpx/s", "value": v, "unit": "Mpx/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": MPX / v * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[2] whole-Antarctic tiled inference, 396 tiles -> 18000x22000 px @250 m",
                       "num_residual_blocks": 12},
            "cpu_baseline": {"value": v, "unit": "Mpx/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
def umma_flops(cin, cout_real, n, h, w):
    return 2.0 * 9 * cin * cout_real * n * h * w


def instrumented_roofline(model, grids, kw, peak_tf):
    """Re-runs a strip of interior tiles with CUDA events (on the launching stream) around every
    tcgen05 conv launch: the persistent trunk kernel (dominant) and the per-layer conv kernel."""
    from deepbedmap_b200 import ops, tiler
    recs = {"umma_trunk_kernel": [], "umma_conv3x3_kernel": []}
    orig = ops.call

    def wrapped(name, *a):
        if name not in ("dbm_trunk_umma", "dbm_conv3x3_umma"):
            return orig(name, *a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, *a)
        e1.record()
        if name == "dbm_trunk_umma":
            n, h, w = a[2], a[3], a[4]
            ws = model._ws[(n, h, w)]
            fl = ws["flops"]
            recs["umma_trunk_kernel"].append((e0, e1, fl))
        else:
            cin, coutp, n, h, w = a[2], a[5], a[6], a[7], a[8]
            cout = 18 if (coutp == 32 and a[12] is None) else coutp   # offset convs: 18 real channels
            recs["umma_conv3x3_kernel"].append((e0, e1, 2.0 * 9 * cin * cout * n * h * w))

    ops.call = wrapped
    try:
        sub = dict(kw)
        sub["final_shape"] = (min(kw["final_shape"][0], 3000), min(kw["final_shape"][1], 6000))
        tiler.predict_continent(model, None, None, None, None, grids=grids, to_host=False, **sub)
        torch.cuda.synchronize()
    finally:
        ops.call = orig
    out = {}
    for k, v in recs.items():
        if not v:
            continue
        ms = sum(a.elapsed_time(b) for a, b, _ in v)
        fl = sum(f for _, _, f in v)
        out[k] = {"tflops": fl / (ms * 1e-3) / 1e12, "launches_timed": len(v), "avg_launch_us": ms * 1e3 / len(v),
                  "flops_per_launch_avg": fl / len(v), "ms_total": ms}
    dom = max(out, key=lambda k: out[k]["ms_total"])
    ach = out[dom]["tflops"]
    # DRAM traffic of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum of one
    # `ncu --set full` capture (profiles/r1h_trunk_kernel_ncu_full_raw.csv: 13.67 + 10.84 GB for a launch of
    # 4 interior tiles = 5.72 TFLOP), scaled to this run's average launch by its algorithmic FLOPs
    traffic = TRUNK_DRAM_BYTES_PER_FLOP * out[dom]["flops_per_launch_avg"] if dom == "umma_trunk_kernel" else None
    return {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
            "traffic": traffic, "traffic_source": "ncu --set full, profiles/r1h_trunk_kernel_ncu_full_raw.csv",
            "kernel": dom, "launches_timed": out[dom]["launches_timed"],
            "avg_launch_us": out[dom]["avg_launch_us"], "flops_per_launch_avg": out[dom]["flops_per_launch_avg"],
            "all_tcgen05_conv_kernels": {k: {kk: vv for kk, vv in v.items() if kk != "ms_total"} | {
                "frac": v["tflops"] / peak_tf} for k, v in out.items()}}


def train_bench(rank, world, steps=20, warmup=3, batch=128):
    from deepbedmap_b200 import train as T
    g, g_opt, d, d_opt = T.compile_srgan_model()
    gen = torch.Generator(device="cuda").manual_seed(42 + rank)
    r = lambda *s: torch.rand(*s, generator=gen, device="cuda")
    arrays = {"X": r(batch, 1, 11, 11), "W1": r(batch, 1, 110, 110), "W2": r(batch, 2, 22, 22), "W3": r(batch, 1, 11, 11),
              "Y": r(batch, 1, 36, 36)}
    # the per-minibatch body of trainer() (srgan_train.py:1286-1308): both steps on the same device batch,
    # the generator step reusing the graph-keeping forward the discriminator step ran (weights unchanged between)
    def eager_step():
        T.train_eval_discriminator(arrays, g, d, d_opt, share_generator_forward=True)
        T.train_eval_generator(arrays, g, d, g_opt)
    launch = "CUDA graph replay (deepbedmap_b200.train.GraphedTrainStep)"
    try:
        # the step (and, data parallel, its bucketed NCCL all-reduces) is captured once as a CUDA graph and replayed
        graphed = T.GraphedTrainStep(arrays, g, g_opt, d, d_opt)
        step = lambda: graphed.step(arrays)
    except Exception as ex:   # reported in the line, never hidden
        launch = f"eager (graph capture failed: {repr(ex)[:120]})"
        step = eager_step
    for _ in range(warmup):
        step()
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    # algorithmic work of a step: 6.4995 GFLOP per sample with the generator forward counted once
    gflop_step = 6.4995 * batch * world
    return {"metric": "train steps/s (D-step + G-step, batch 128 per GPU)", "value": steps / (ms * 1e-3),
            "algorithmic_tflops": gflop_step * steps / (ms * 1e-3) / 1e3,
            "unit": "steps/s", "ms_per_step": ms / steps, "steps": steps, "warmup": warmup, "scaling": "weak",
            "global_batch": batch * world, "dtype": "bf16 tensor-core operands / fp32 accumulation for every 3x3 conv of G and D (fwd, dgrad, wgrad) and the "
                                               "deformable contraction; fp32 stem, bilinear sampling, BN, losses, Adam",
            "generator_forward": "one per step, shared by the D-step and the G-step (the reference runs it twice "
                                 "with unchanged weights; SURVEY 8d counts it once)",
            "launch": launch,
            "allreduce": "bucketed, launched from inside backward (NCCL, overlapped)" if world > 1 else "none (1 GPU)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-tiles", type=int, default=4)
    ap.add_argument("--cpu-crop", type=int, default=192)
    ap.add_argument("--scale", type=float, default=1.0, help="debug only: shrink the continent (invalid as a result)")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    rank, world, local = dist_setup(args.gpus)
    from deepbedmap_b200 import GeneratorModel, _lib, tiler
    peak_tf, peak_hbm, peak_src = measured_peaks()

    if args.scale != 1.0:
        fy = max(1000, int(FULL["final_shape"][0] * args.scale) // 1000 * 1000)
        fx = max(1000, int(FULL["final_shape"][1] * args.scale) // 1000 * 1000)
        final_shape, grid = (fy, fx), (fy // 4 + 2, fx // 4 + 2)
    else:
        final_shape, grid = FULL["final_shape"], FULL["grid"]
    mpx = final_shape[0] * final_shape[1] / 1e6
    kw = dict(final_shape=final_shape, ary_shape=(1000, 1000), stride=(1000, 1000), xtrapad=(18, 18),
              batch_tiles=args.batch_tiles)

    model = GeneratorModel(num_residual_blocks=12, residual_scaling=0.1, precision="bf16", seed=0)
    X, W1, W2, W3 = synth_grids_device(grid)
    grids = tiler.ContinentGrids(X, W1, W2, W3)

    def step():
        return tiler.predict_continent(model, None, None, None, None, grids=grids, to_host=False, **kw)

    for _ in range(args.warmup):
        step()
    barrier(world)
    l0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier(world)
    launches = _lib.launch_count - l0
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    value = mpx * args.steps / (ms * 1e-3)

    # every rank runs the instrumented strip (predict_continent ends in a collective gather); rank 0's
    # CUDA-event timings are the ones reported
    roof = instrumented_roofline(model, grids, kw, peak_tf)
    barrier(world)

    # ---- end to end: pinned host grids -> host DEM ----
    e2e = None
    if not args.no_e2e:
        host = [torch.empty(t.shape, dtype=torch.float32, pin_memory=True).copy_(t) for t in (X, W1, W2, W3)]
        torch.cuda.synchronize()
        del grids, X, W1, W2, W3
        torch.cuda.empty_cache()
        h2d = sum(t.numel() * 4 for t in host)
        out_host = torch.empty(1, final_shape[0], final_shape[1], dtype=torch.float32, pin_memory=True) if rank == 0 else None
        tiler.predict_continent(model, *host, out=out_host, **kw)  # warm-up
        barrier(world)
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 2))
        for _ in range(n_e2e):
            out = tiler.predict_continent(model, *host, out=out_host, **kw)
        barrier(world)
        dt = max_over_ranks((time.perf_counter() - t0) * 1e3, world)
        e2e = {"value": mpx * n_e2e / (dt * 1e-3), "unit": "Mpx/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(final_shape[0] * final_shape[1] * 4), "steps": n_e2e,
               "api": "deepbedmap_b200.predict_continent(model, X, W1, W2, W3, out=pinned) with pinned host arrays; "
                      "band-wise H2D overlapped with compute"}
        if rank == 0:
            assert out is not None and out.shape == (1, final_shape[0], final_shape[1])
        del host, out

    train = None
    if not args.no_train:
        try:
            train = train_bench(rank, world)
        except Exception as ex:  # reported, never hidden
            train = {"error": repr(ex)[:300]}

    if rank == 0:
        cpu = None
        if not args.no_cpu:
            v, cores, sample = cpu_reference_rate(crop=args.cpu_crop)
            cpu = {"value": v, "unit": "Mpx/s", "cores": cores, "kind": "port", "sample": sample}
        line = {"metric": "continent inference output Mpx/s", "value": value, "unit": "Mpx/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16 (fp32 accumulate, fp32 residual stream)",
                "data": "synthetic",
                "config": {"workload": "configs[2] whole-Antarctic tiled inference, 396 tiles -> 18000x22000 px @250 m"
                           if args.scale == 1.0 else f"DEBUG scaled continent {final_shape}",
                           "num_residual_blocks": 12, "batch_tiles": args.batch_tiles, "tile_split": f"contiguous/{world}",
                           "l2": "inputs (10.9 GB) and per-layer activations exceed the 126 MB L2; no flush needed"},
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roof,
                "roofline_peak_source": peak_src, "cpu_baseline": cpu, "train": train}
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
