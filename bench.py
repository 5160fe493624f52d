#!/usr/bin/env python
"""bench.py — headline benchmark of the finufft_b200 hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload c3_t1|c3_t2|c2_t2|c2_t1|c4_t1|c1_t1] [--M pts] [--dist uniform|cluster]

Workload (default c3_t1 = BASELINE.json configs[2], the config the metric is quoted on and
that fits one GPU): 3D type 1, single precision, 256^3 modes (fine grid 512^3), M = 1e8
uniform-random points, tol 1e-6, sigma 2.  One "step" = one execute of the plan
(spread + FFT + deconvolve) on a fresh strength vector; setpts is done once before and
timed separately (reported as setpts_ms).

value   NU points/s over all GPUs, device-resident inputs, CUDA-event timing, max over ranks.
e2e     same metric through the host-pointer C ABI (finufft[f]_execute) with pinned host
        buffers: H2D of the strengths and D2H of the modes are inside the timed region.
roofline  the spread (type 1) / interp (type 2) kernel: algorithmic bytes of SURVEY.md 8(d)
        per launch / its average CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline  the CPU oracle port (oracle/liboracle.so, OpenMP) on a bounded sample, rank 0.

N>1: launched under torchrun, one process per GPU; each rank runs the same transform on its
own strength vector (the "batched ntransf split across GPUs" sharding of SURVEY.md 8(e): no
data-path collective), so scaling is weak.  Inputs (0.8 GB strengths, 1 GB grid) exceed the
126 MB L2, so no explicit L2 flush is needed between timed iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (type, modes, M, tol, dtype, ntr)
    "c3_t1": (1, (256, 256, 256), 100_000_000, 1e-6, "complex64", 1),
    "c3_t2": (2, (256, 256, 256), 100_000_000, 1e-6, "complex64", 1),
    "c2_t2": (2, (2048, 2048), 100_000_000, 1e-5, "complex64", 1),
    "c2_t1": (1, (2048, 2048), 100_000_000, 1e-5, "complex64", 1),
    "c4_t1": (1, (512, 512), 10_000_000, 1e-9, "complex128", 8),
    # configs[0], the reference's own CPU-runnable case (perftest --prec d --type 1 --N1 1e6 --M 1e7)
    "c1_t1": (1, (1_000_000,), 10_000_000, 1e-9, "complex128", 1),
}
# type 3 (configs[4]): M sources, N = M targets with frequencies of half-width 107.5 per dim
# (perftest/perftest.cpp:197-202 with N1=N2=N3=215); benched by tools/bench_type3.py
METRIC = "NU points/sec (spread+FFT+deconv)"


def algorithmic_bytes(dim, M, nf, real_bytes):
    """spread/interp = (4 + d*s + 2s)*M + 2s*G."""
    G = int(np.prod(nf))
    return (4 + dim * real_bytes + 2 * real_bytes) * M + 2 * real_bytes * G


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def ncu_traffic(workload, dist):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/traffic.json, written by tools/ncu_summary.py); None if no capture is recorded."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(f"{workload}:{dist}", {}).get("dram_bytes")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap,power.draw")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                 "--format=csv,noheader,nounits", "-lms", "25"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_points(dim, M, rt, seed, dist, nf):
    """Synthetic nonuniform points: uniform in [-pi,pi)^d, or clustered = all points inside an
    8-cell cube of the fine grid (SURVEY.md 8(d) definition)."""
    rng = np.random.default_rng(seed)
    out = []
    for d in range(dim):
        if dist == "cluster":
            h = 2 * np.pi / nf[d]
            out.append(rng.uniform(0, 8 * h, M).astype(rt))
        else:
            out.append(rng.uniform(-np.pi, np.pi, M).astype(rt))
    return out


# --------------------------------------------------------------------------------- CPU arm
def cpu_run(type_, modes, M, tol, dtype, steps, warmup, dist):
    """Time the CPU oracle port on M points (all host threads)."""
    from oracle import build as obuild
    obuild.build_oracle()
    from oracle import oracle as O
    rt = np.float32 if dtype == "complex64" else np.float64
    dim = len(modes)
    # every core this process may run on: torchrun exports OMP_NUM_THREADS=1, which would
    # otherwise turn the "all host threads" baseline into a single-thread one
    try:
        nthr = len(os.sched_getaffinity(0))
    except AttributeError:
        nthr = os.cpu_count() or 1
    nthr = max(nthr, O.max_threads())
    plan = O.Plan(type_, list(modes), 1, 1, tol, rt, sigma=2.0, nthr=nthr)
    pts = synth_points(dim, M, rt, 1234, dist, plan.nf) + [None] * (3 - dim)
    t0 = time.perf_counter()
    plan.setpts(*pts)
    t_setpts = time.perf_counter() - t0
    rng = np.random.default_rng(5)
    n_in = M if type_ == 1 else int(np.prod(modes))
    data = (rng.standard_normal(n_in) + 1j * rng.standard_normal(n_in)).astype(dtype)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        plan.execute(data)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    plan.destroy()
    return float(np.mean(times)), t_setpts, nthr


# --------------------------------------------------------------------------------- z-slab leg
def zslab_leg(type_, modes, M, tol, rank, world, dev, steps, warmup, barrier):
    """Strong scaling of ONE transform of M points over `world` GPUs: z-slab decomposition of the
    fine grid (finufft_b200/zslab.py: ghost-plane send/recv + slab->pencil all_to_all over NCCL).
    Points are generated already partitioned by slab; outputs stay sharded."""
    import torch
    import torch.distributed as dist_
    from finufft_b200.zslab import SlabPlan
    try:
        sp = SlabPlan(type_, modes, tol, 1 if type_ == 1 else -1, "complex64")
    except ValueError as exc:
        return {"unavailable": str(exc)}
    Ml = M // world
    g = torch.Generator(device=dev)
    g.manual_seed(500 + rank)
    h = 2 * np.pi / sp.nf[0]
    top = float(np.nextafter(np.float32(-np.pi + h * sp.z1), np.float32(-4)))
    z = (-np.pi + h * (sp.z0 + sp.nz * torch.rand(Ml, device=dev, generator=g))).float().clamp_(max=top)
    y = ((torch.rand(Ml, device=dev, generator=g) * 2 - 1) * np.pi).float()
    x = ((torch.rand(Ml, device=dev, generator=g) * 2 - 1) * np.pi).float()
    sp.setpts(z, y, x)
    shape = (Ml,) if type_ == 1 else (modes[0], sp.y_hi - sp.y_lo, modes[2])
    data = torch.view_as_complex(torch.randn(shape + (2,), device=dev, generator=g))
    for _ in range(max(warmup, 3)):
        sp.execute(data)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sp.execute(data)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist_.all_reduce(t, op=dist_.ReduceOp.MAX)
    sp.destroy()
    ms = float(t.item())
    return {"value": M / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms, "scaling": "strong",
            "workload": f"one transform, M={M:.3g} total points pre-partitioned by z-slab over "
                        f"{world} GPUs, outputs left sharded"}


# --------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3_t1", choices=sorted(WORKLOADS))
    ap.add_argument("--M", type=float, default=None, help="override number of points")
    ap.add_argument("--dist", default="uniform", choices=["uniform", "cluster"])
    ap.add_argument("--cpu-sample", type=float, default=None,
                    help="points in the CPU sample (default: sized for ~10-30 s)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-zslab", action="store_true",
                    help="N>1: skip the extra strong-scaling (z-slab sharded single transform) leg")
    args = ap.parse_args()

    type_, modes, M, tol, dtype, ntr = WORKLOADS[args.workload]
    if args.M:
        M = int(args.M)
    dim = len(modes)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    rt = np.float32 if dtype == "complex64" else np.float64
    rbytes = 4 if rt == np.float32 else 8
    config = {"workload": f"{dim}D type {type_} {'f32' if rbytes == 4 else 'f64'}, modes "
              + "x".join(map(str, modes)) + f", M={M:.3g} {args.dist} points, tol={tol:g}, "
              f"sigma=2, ntransf={ntr}" + (f" per GPU x {world} GPUs" if world > 1 else ""),
              "name": args.workload, "l2": "inputs larger than L2 (no flush needed)",
              "sharding": "independent transforms per GPU, no collective" if world > 1 else "1 GPU"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        cpu_M = int(args.cpu_sample or min(M, 10_000_000))
        sec, t_setpts, nthr = cpu_run(type_, modes, cpu_M, tol, dtype, args.steps,
                                      max(args.warmup, 0), args.dist)
        val = cpu_M / sec
        line = {"metric": METRIC, "value": val, "unit": "points/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if rbytes == 4 else "f64", "data": "synthetic", "config": config,
                "impl": "reference",
                "cpu_baseline": {"value": val, "unit": "points/s", "cores": nthr, "kind": "port",
                                 "sample": f"same grid and tolerance, {cpu_M:.3g} of the "
                                           f"{M:.3g} points per step (oracle/liboracle.so, "
                                           "OpenMP; the reference CPU library itself cannot be "
                                           "built offline)"},
                "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist_
    import finufft_b200 as F

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist_.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist_.barrier()
        torch.cuda.synchronize()

    plan = F.Plan(type_, modes, ntr, tol, 1, dtype, upsampfac=2.0, gpu_device_id=local)
    info = plan.info()
    nf = info["nf"]
    pts_h = synth_points(dim, M, rt, 1234 + rank, args.dist, nf[::-1])
    pts = [torch.from_numpy(p).to(dev) for p in pts_h]
    plan.enable_profiling(True)
    plan.setpts(*pts)  # untimed: first-call module loading, pool growth
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plan.setpts(*pts)
    torch.cuda.synchronize()
    setpts_wall_ms = (time.perf_counter() - t0) * 1e3
    setpts_ms = plan.stage_ms()["setpts"]
    info = plan.info()

    cdt = torch.complex64 if dtype == "complex64" else torch.complex128
    gen = torch.Generator(device=dev)
    gen.manual_seed(99 + rank)
    if type_ == 1:
        in_shape = (ntr, M) if ntr > 1 else (M,)
    else:
        in_shape = ((ntr,) + tuple(modes)) if ntr > 1 else tuple(modes)
    rdt = torch.float32 if rbytes == 4 else torch.float64
    data = torch.view_as_complex(torch.randn(in_shape + (2,), dtype=rdt, device=dev,
                                             generator=gen))
    out = plan.execute(data)
    torch.cuda.synchronize()

    for _ in range(args.warmup):
        plan.execute(data, out)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = plan.launch_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    stage = {"spreadinterp": [], "fft": [], "deconv": []}
    barrier()
    e0.record()
    for _ in range(args.steps):
        plan.execute(data, out)
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    launches = plan.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    # per-stage durations: separate pass so event queries do not perturb the timed loop
    for _ in range(max(3, min(args.steps, 10))):
        plan.execute(data, out)
        s = plan.stage_ms()
        for k in stage:
            stage[k].append(s[k])
    stage_avg = {k: float(np.mean(v)) for k, v in stage.items()}

    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist_.all_reduce(t, op=dist_.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * M * ntr / (ms_per_step * 1e-3)

    # ------------------------------------------------------------------ e2e: host-pointer ABI
    hplan = F.HostPlan(type_, modes, ntr, tol, 1, dtype, upsampfac=2.0, allow_eps_too_small=1)
    hplan.setpts(*pts_h)
    n_in = int(np.prod(in_shape))
    n_out = ntr * (int(np.prod(modes)) if type_ == 1 else M)
    h_in = torch.empty(in_shape, dtype=cdt).pin_memory()
    h_in.copy_(data.cpu())
    out_shape = (((ntr,) if ntr > 1 else ()) + (tuple(modes) if type_ == 1 else (M,)))
    h_out = torch.empty(out_shape, dtype=cdt).pin_memory()
    h_in_np, h_out_np = h_in.numpy(), h_out.numpy()
    e2e_steps = max(2, min(args.steps, 5))
    hplan.execute(h_in_np, out=h_out_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        hplan.execute(h_in_np, out=h_out_np)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist_.all_reduce(te, op=dist_.ReduceOp.MAX)
    e2e_val = world * M * ntr / (float(te.item()) * 1e-3)
    cbytes = 8 if rbytes == 4 else 16
    hplan.destroy()

    zslab = None
    if world > 1 and dim == 3 and dtype == "complex64" and ntr == 1 and not args.no_zslab:
        zslab = zslab_leg(type_, modes, M, tol, rank, world, dev, args.steps, args.warmup, barrier)

    if rank == 0:
        peak, peak_kind = measured_hbm_peak()
        abytes = algorithmic_bytes(dim, M, nf, rbytes)
        kernel_ms = stage_avg["spreadinterp"] / ntr  # one launch per transform
        sweep_on = os.environ.get("B200_NUFFT_SWEEP", "1") != "0"
        swept = sweep_on and dim == 3 and rbytes == 4 and info["ns"] in (6, 7)
        tf = "true" if type_ == 1 else "false"
        if swept:
            kernel_name = f"k_sweep3<{info['ns']},{tf}>"
        elif sweep_on and dim == 2:
            kernel_name = f"k_sweep2<{'float' if rbytes == 4 else 'double'},{info['ns']},{tf}>"
        else:
            kernel_name = "k_spread" if type_ == 1 else "k_interp"
        achieved = abytes / (kernel_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": kernel_name,
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic(args.workload, args.dist), "peak_source": peak_kind,
                    "algorithmic_bytes": abytes, "kernel_ms": kernel_ms}
        # secondary roof (SURVEY.md 8(d) "which roof binds"): ns^d complex cell updates per
        # point against the packed-FFMA2 rate measured on this part (tools/micro/micro1.cu,
        # profiles/r1_micro_ffma2_red.txt: 17.6e12 cell updates/s)
        if rbytes == 4:
            cells = float(info["ns"]) ** dim * M
            roofline["fp32"] = {"achieved_tcell_s": cells / (kernel_ms * 1e-3) / 1e12,
                                "peak_tcell_s": 17.6, "frac": cells / (kernel_ms * 1e-3) / 17.6e12,
                                "floor_ms": cells / 17.6e12 * 1e3}
        cpu = None
        if not args.no_cpu:
            cpu_M = int(args.cpu_sample or min(M, 10_000_000))
            try:
                sec, _, nthr = cpu_run(type_, modes, cpu_M, tol, dtype, 1, 0, args.dist)
                cpu = {"value": cpu_M / sec, "unit": "points/s", "cores": nthr, "kind": "port",
                       "sample": f"same grid and tolerance, {cpu_M:.3g} of the {M:.3g} points, "
                                 "1 execute (oracle/liboracle.so, OpenMP)"}
            except Exception as exc:  # the checker failing must not lose the GPU number
                cpu = {"value": None, "unit": "points/s", "cores": 0, "kind": "port",
                       "sample": f"failed: {exc}"}
        line = {"metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if rbytes == 4 else "f64", "data": "synthetic", "config": config,
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "points/s",
                        "h2d_bytes_per_step": n_in * cbytes, "d2h_bytes_per_step": n_out * cbytes,
                        "ms_per_step": float(te.item()),
                        "api": "finufft[f]_execute (host pointers, pinned buffers)"},
                "gpu_launches": launches,
                "roofline": roofline, "cpu_baseline": cpu,
                "stages_ms": stage_avg, "setpts_ms": setpts_ms, "setpts_wall_ms": setpts_wall_ms,
                "value_with_setpts": world * M * ntr / ((ms_per_step + setpts_ms) * 1e-3),
                "plan": {"ns": info["ns"], "nf": nf, "nsub": info["nsub"]}}
        if zslab is not None:
            line["zslab"] = zslab
        print(json.dumps(line), flush=True)
    plan.destroy()
    if world > 1:
        dist_.destroy_process_group()


if __name__ == "__main__":
    main()
