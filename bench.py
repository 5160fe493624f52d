#!/usr/bin/env python
"""bench.py — headline benchmark of the finufft_b200 hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload c3_t1|c3_t2|c2_t2|c2_t1|c4_t1|c1_t1] [--M pts] [--dist uniform|cluster]

Workload (default c3_t1 = BASELINE.json configs[2], the config the metric is quoted on and that
fits one GPU): 3D type 1, single precision, 256^3 modes (fine grid 512^3), M = 1e8 points,
tol 1e-6, sigma 2.  Inputs are the reference perftest's own synthetic data (tools/perfdata.py =
perftest/randunif.h streams).  One "step" = one execute of the plan (spread + FFT + deconvolve);
two strength vectors alternate between steps; setpts is done once before and timed separately
(setpts_ms).

value   NU points/s of the whole job, device-resident inputs, CUDA-event timing, max over ranks.
e2e     same metric with HOST buffers: N = 1 through finufft[f]_execute (host-pointer C ABI,
        pinned buffers, H2D of the strengths and D2H of the modes inside the timed region);
        N > 1 pinned host -> device copy of the rank's strengths + sharded execute + D2H of the
        rank's mode block.
roofline  the spread (type 1) / interp (type 2) kernel: algorithmic bytes of SURVEY.md 8(d) per
        launch / its average CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
accuracy  the OUTPUT OF THE TIMED PLAN against a float64 direct sum at a few modes / points,
        computed on the GPU in the same process (asserted below a sanity bound; the parity
        gates proper are tests/test_gpu_baseline_sizes.py).
cpu_baseline  the CPU arm (see --impl reference) on a bounded sample, rank 0, N = 1.

N > 1 (torchrun, one process per GPU):
  3D single-vector workloads (c3_*)  ONE transform of M points sharded over the N GPUs by z-slabs
        of the fine grid (include/b200_sharded.h): every rank holds M/N of the points (an
        arbitrary share: the library routes them), the ghost-plane exchange and the slab<->pencil
        transpose are NCCL calls inside the step, outputs stay sharded ("scaling": "strong").
        Side keys: the same with points pre-partitioned by slab, the cost of gathering the modes,
        the per-stage breakdown, and N independent replicas (weak).
  batched workloads (c4_t1, ntransf = 64)  the vectors split across the GPUs, no collective
        ("strong": the 64 vectors are the fixed total).
  everything else: replicas only (weak).
Inputs (0.8 GB strengths, 1 GB grid) exceed the 126 MB L2, so no explicit flush is needed.
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
sys.path.insert(0, os.path.join(ROOT, "tools"))

WORKLOADS = {
    # name: (type, modes, M, tol, dtype, ntr)
    "c3_t1": (1, (256, 256, 256), 100_000_000, 1e-6, "complex64", 1),
    "c3_t2": (2, (256, 256, 256), 100_000_000, 1e-6, "complex64", 1),
    "c2_t2": (2, (2048, 2048), 100_000_000, 1e-5, "complex64", 1),
    "c2_t1": (1, (2048, 2048), 100_000_000, 1e-5, "complex64", 1),
    "c4_t1": (1, (512, 512), 10_000_000, 1e-9, "complex128", 64),
    # configs[0], the reference's own CPU-runnable case (perftest --prec d --type 1 --N1 1e6 --M 1e7)
    "c1_t1": (1, (1_000_000,), 10_000_000, 1e-9, "complex128", 1),
}
# type 3 (configs[4]): M sources, N = M targets with frequencies of half-width 107.5 per dim
# (perftest/perftest.cpp:197-202 with N1=N2=N3=215); benched by tools/bench_type3.py
METRIC = "NU points/sec (spread+FFT+deconv)"


def algorithmic_bytes(dim, M, cells, real_bytes):
    """spread/interp = (4 + d*s + 2s)*M + 2s*G."""
    return (4 + dim * real_bytes + 2 * real_bytes) * M + 2 * real_bytes * int(cells)


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
    (profiles/traffic.json, written by tools/ncu_summary.py); None if no capture is recorded.
    It is NOT measured in this run (a run under ncu is never a bench value)."""
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


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------- CPU arm
def cpu_sample_shape(modes, M, ntr):
    """Bounded sample of the workload for the CPU arm: 1/8 of the points on a grid with 1/8 of the
    modes (3D: every dimension halved), i.e. the SAME points-per-cell density and the same
    spread : FFT cost ratio as the full workload; 2D / 1D workloads keep their grid and take
    1e7 points; one vector of a batched workload."""
    if len(modes) == 3 and M > 12_500_000:
        return tuple(m // 2 for m in modes), M // 8, "1/8 of the points on a grid of half the modes per dimension (same points per cell)"
    if M > 10_000_000:
        return tuple(modes), 10_000_000, "1e7 of the points on the full grid"
    return tuple(modes), M, "all points, full grid" + (", one of the vectors" if ntr > 1 else "")


def cpu_run(type_, modes, M, tol, dtype, steps, warmup, dist):
    """Time the CPU arm on M points with all host threads.  The reference CPU library itself
    compiled here (oracle/_ref/libfinufft_ref.so, its own spreader / sort / deconvolve with
    OpenMP) when that build exists, else the scalar restatement oracle/liboracle.so."""
    from oracle import build as obuild
    obuild.build_oracle()
    from oracle import oracle as O
    import perfdata
    rt = np.float32 if dtype == "complex64" else np.float64
    dim = len(modes)
    # every core this process may run on: torchrun exports OMP_NUM_THREADS=1, which would
    # otherwise turn the "all host threads" baseline into a single-thread one
    nthr = max(host_threads(), O.max_threads())
    kind = "port"
    make = O.Plan
    if getattr(O, "have_reference", lambda: False)():
        kind, make = "reference", O.RefPlan
    plan = make(type_, list(modes), 1, 1, tol, rt, sigma=2.0, nthr=nthr)
    pts = perfdata.points(dim, M, rt, dist, plan.nf) + [None] * (3 - dim)
    t0 = time.perf_counter()
    plan.setpts(*pts)
    t_setpts = time.perf_counter() - t0
    n_in = M if type_ == 1 else int(np.prod(modes))
    data = perfdata.strengths(n_in, dtype, "C" if type_ == 1 else "FK")
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        plan.execute(data)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    plan.destroy()
    return float(np.mean(times)), t_setpts, nthr, kind


# --------------------------------------------------------------------------------- accuracy
def direct_sum_check(torch, type_, modes, isign, pts_dev, c_dev, fk_dev, reduce=None, nprobe=6):
    """rel-l2 error of the plan's own output at a few modes (type 1) or points (type 2) against a
    float64 direct sum on the GPU.  pts_dev = [x, y, z][:dim] (library order), modes python
    order (slowest first), fk_dev the full mode array.  reduce = all-reduce(sum) for sharded
    points (type 1 only)."""
    dim = len(modes)
    lib_modes = modes[::-1]  # x first
    rng = np.random.default_rng(7)
    if type_ == 1:
        ks = [[int(rng.integers(-(m // 2), (m - 1) // 2 + 1)) for m in lib_modes]
              for _ in range(nprobe - 2)]
        ks += [[-(m // 2) for m in lib_modes], [(m - 1) // 2 for m in lib_modes]]
        M = pts_dev[0].numel()
        acc = torch.zeros(len(ks), dtype=torch.complex128, device=c_dev.device)
        step = 10_000_000
        for a in range(0, M, step):
            b = min(M, a + step)
            cc = c_dev[a:b].to(torch.complex128)
            for i, k in enumerate(ks):
                ph = torch.zeros(b - a, dtype=torch.float64, device=c_dev.device)
                for d in range(dim):
                    ph += k[d] * pts_dev[d][a:b].double()
                acc[i] += torch.sum(cc * torch.polar(torch.ones_like(ph), isign * ph))
        if reduce is not None:
            acc = reduce(acc)
        got = torch.stack([fk_dev[tuple(k[d] + lib_modes[d] // 2 for d in range(dim))[::-1]]
                           for k in ks]).to(torch.complex128)
    else:
        js = list(range(nprobe))
        f = fk_dev.to(torch.complex128)
        acc = []
        for j in js:
            t = f
            for d in range(dim):  # contract the fastest (last) axis = library dimension d
                m = lib_modes[d]
                k = torch.arange(-(m // 2), (m - 1) // 2 + 1, dtype=torch.float64, device=f.device)
                ph = k * pts_dev[d][j].double()
                t = t @ torch.polar(torch.ones_like(ph), isign * ph)
            acc.append(t)
        acc = torch.stack(acc)
        got = c_dev[:nprobe].to(torch.complex128)
    err = float(torch.linalg.norm(got - acc) / torch.linalg.norm(acc))
    return {"relerr": err, "against": f"float64 direct sum on the GPU, {len(acc)} "
            + ("modes" if type_ == 1 else "points") + " of the timed plan's output"}


# --------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3_t1", choices=sorted(WORKLOADS))
    ap.add_argument("--M", type=float, default=None, help="override number of points")
    ap.add_argument("--ntr", type=int, default=None, help="override the number of vectors")
    ap.add_argument("--dist", default="uniform", choices=["uniform", "cluster"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the side legs (other type, replicas, pre-partitioned points)")
    args = ap.parse_args()

    type_, modes, M, tol, dtype, ntr = WORKLOADS[args.workload]
    if args.M:
        M = int(args.M)
    if args.ntr:
        ntr = args.ntr
    dim = len(modes)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    rt = np.float32 if dtype == "complex64" else np.float64
    rbytes = 4 if rt == np.float32 else 8
    prec = "f32" if rbytes == 4 else "f64"
    sharded = world > 1 and dim == 3 and ntr == 1
    batched = world > 1 and ntr > 1
    if sharded:
        if args.dist == "cluster":
            sharding = (f"one transform sharded over {world} GPUs; every rank holds M/N of the "
                        "(clustered) points, the library replicates the occupied window of the "
                        "grid and reduces it onto the owning slabs; outputs sharded")
        else:
            sharding = (f"one transform sharded over {world} GPUs by z-slabs of the fine grid; "
                        "inputs resident in the HBM of the GPU that owns them (every rank holds "
                        "the M/N points of its slab), both exchanges inside the step, outputs "
                        "sharded; `arbitrary_points` = same with every rank holding an arbitrary "
                        "M/N share that the library routes inside the step")
    elif batched:
        sharding = f"the {ntr} vectors split across {world} GPUs, no collective"
    elif world > 1:
        sharding = "replicas only: independent transforms per GPU, no collective"
    else:
        sharding = "1 GPU"
    config = {"workload": f"{dim}D type {type_} {prec}, modes " + "x".join(map(str, modes))
              + f", M={M:.3g} {args.dist} points, tol={tol:g}, sigma=2, ntransf={ntr}",
              "name": args.workload, "l2": "inputs larger than L2 (no flush needed)",
              "inputs": "perftest/randunif.h streams (tools/perfdata.py)",
              "sharding": sharding}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        cmodes, cM, what = cpu_sample_shape(modes, M, ntr)
        sec, t_setpts, nthr, kind = cpu_run(type_, cmodes, cM, tol, dtype, args.steps,
                                            max(args.warmup, 0), args.dist)
        val = cM / sec
        label = ("the reference's own CPU code (include/finufft/*.hpp, src/*.cpp) compiled "
                 "where it lies with stand-ins for its absent third-party SIMD and FFT "
                 "libraries" if kind == "reference" else
                 "scalar restatement oracle/liboracle.so with its own FFT, OpenMP - not the "
                 "FINUFFT binary")
        line = {"metric": METRIC, "value": val, "unit": "points/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
                "higher_is_better": True, "scaling": "strong" if (sharded or batched) else "weak",
                "vs_baseline": None, "dtype": prec, "data": "synthetic", "config": config,
                "impl": "reference",
                "cpu_baseline": {"value": val, "unit": "points/s", "cores": nthr, "kind": kind,
                                 "sample": f"{what}: modes {'x'.join(map(str, cmodes))}, "
                                           f"M={cM:.3g} per step; {label}",
                                 "setpts_s": t_setpts},
                "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist_
    import finufft_b200 as F
    import perfdata

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist_.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist_.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist_.all_reduce(t, op=dist_.ReduceOp.MAX)
        return float(t.item())

    cdt = torch.complex64 if dtype == "complex64" else torch.complex128
    rdt = torch.float32 if rbytes == 4 else torch.float64
    cbytes = 2 * rbytes
    isign = 1
    peak, peak_kind = measured_hbm_peak()
    nmodes = int(np.prod(modes))

    def timed(fn, steps, warmup):
        """(ms per step, max over ranks) of fn(i) with CUDA events between barriers."""
        for i in range(warmup):
            fn(i)
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        return allmax(e0.elapsed_time(e1)) / steps

    def timed_wall(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            fn(i)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        return allmax(ms)

    sampler = ClockSampler(local)
    line = None

    if sharded:
        # ============================================================== one transform, N GPUs
        from finufft_b200.sharded import ShardedPlan
        lo, hi = rank * M // world, (rank + 1) * M // world
        Ml = hi - lo
        sp = ShardedPlan(type_, modes, eps=tol, isign=isign, dtype=dtype, upsampfac=2.0)
        info = sp.info()
        nf = info["nf"]
        pts_h = perfdata.points(3, Ml, rt, args.dist, nf, first=lo)        # x, y, z share
        x, y, z_any = (torch.from_numpy(p).to(dev) for p in pts_h)
        resident = args.dist != "cluster"
        if resident:
            # the same streams, the z stream mapped into this rank's slab: uniform points, every
            # one of them owned by this rank (a margin of 1e-5 slab widths keeps the single-
            # precision fold on the right side of the slab faces)
            u = np.empty(Ml, dtype=rt)
            perfdata.fill(u, "Z", 1.0, 0.0, first=lo)                     # U(-1, 1)
            frac = (0.5 * (u.astype(np.float64) + 1.0)) * (1 - 2e-5) + 1e-5
            zres = (-np.pi + (2 * np.pi / nf[2]) * (info["z0"] + info["nz"] * frac)).astype(rt)
            z = torch.from_numpy(zres).to(dev)
        else:
            z = z_any

        def set_points(zz, routed):
            sp.setpts(zz, y, x, routed=routed)          # untimed: module loading, pools, mappings
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            sp.setpts(zz, y, x, routed=routed)
            torch.cuda.synchronize()
            return (allmax((time.perf_counter() - t0) * 1e3), allmax(sp.stage_ms()["setpts"]))
        setpts_wall_ms, setpts_ms = set_points(z, resident)
        info = sp.info()
        if type_ == 1:
            h_in = [torch.from_numpy(perfdata.strengths(Ml, dtype, "C", first=lo)).pin_memory(),
                    torch.from_numpy(perfdata.strengths(Ml, dtype, "FK", first=lo)).pin_memory()]
        else:
            full = [perfdata.strengths(nmodes, dtype, s).reshape(modes) for s in ("FK", "C")]
            h_in = [torch.from_numpy(np.ascontiguousarray(f[:, info["ylo"]:info["yhi"], :]))
                    .pin_memory() for f in full]
        data = [h.to(dev) for h in h_in]
        out = sp.execute(data[0])
        if rank == 0:
            sampler.start()
        l0 = sp.launch_count()
        ms_per_step = timed(lambda i: sp.execute(data[i & 1], out), args.steps, args.warmup)
        launches = (sp.launch_count() - l0) // (args.steps + args.warmup)
        clocks = sampler.stop() if rank == 0 else None
        value = M / (ms_per_step * 1e-3)
        # stage breakdown (max over ranks per stage), separate pass
        stage = {}
        for i in range(5):
            barrier()  # the first exchange of a step would otherwise absorb the ranks' skew
            sp.execute(data[i & 1], out)
            for k, v in sp.stage_ms().items():
                stage.setdefault(k, []).append(v)
        stage_avg = {k: allmax(float(np.mean(v))) for k, v in stage.items() if k != "setpts"}
        # accuracy of the timed plan
        sp.execute(data[0], out)
        if type_ == 1:
            fk_full = sp.gather_modes(out)

            def red(a):
                dist_.all_reduce(a)
                return a
            acc = direct_sum_check(torch, 1, modes, isign, [x, y, z], data[0], fk_full, red)
        else:
            fk_full = torch.from_numpy(full[0]).to(dev)
            acc = direct_sum_check(torch, 2, modes, isign, [x, y, z], out, fk_full)
            acc["relerr"] = allmax(acc["relerr"])
        # e2e: pinned host strengths -> device -> sharded execute -> host block
        h_out = torch.empty(out.shape, dtype=cdt).pin_memory()
        d_in = torch.empty_like(data[0])

        def e2e_step(i):
            d_in.copy_(h_in[i & 1], non_blocking=True)
            sp.execute(d_in, out)
            h_out.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e2e_ms = timed_wall(e2e_step, args.steps, 2)
        extras = {}
        if not args.no_extras:
            # gather of the modes onto every rank (type 1) inside the step
            if type_ == 1:
                ms_g = timed(lambda i: sp.gather_modes(sp.execute(data[i & 1], out)),
                             max(5, args.steps // 2), 2)
                extras["with_mode_gather"] = {"ms_per_step": ms_g, "value": M / (ms_g * 1e-3)}
            # every rank holding an ARBITRARY M/N share of the points: the library routes the
            # coordinates at setpts and the strengths / values inside every step
            if resident:
                wall_any, setpts_any = set_points(z_any, False)
                ms_r = timed(lambda i: sp.execute(data[i & 1], out), max(5, args.steps // 2), 2)
                st_any = {}
                for i in range(3):
                    barrier()
                    sp.execute(data[i & 1], out)
                    for k, v in sp.stage_ms().items():
                        st_any.setdefault(k, []).append(v)
                extras["arbitrary_points"] = {
                    "ms_per_step": ms_r, "value": M / (ms_r * 1e-3), "unit": "points/s",
                    "route_values_ms": allmax(float(np.mean(st_any["route_values"]))),
                    "setpts_ms": setpts_any,
                    "note": "strengths / values routed between the GPUs inside the timed step"}
        win_cells = info["win_n"] * nf[0] * nf[1]
        abytes = algorithmic_bytes(3, info["M_local"], win_cells, rbytes)
        kernel_ms = stage_avg["spreadinterp"]
        kname = f"k_sweep3<{info['ns']},{'true' if type_ == 1 else 'false'}>"
        sp.destroy()
        replicas = None
        if not args.no_extras:
            # N independent transforms of M points each (weak scaling, no collective)
            plan = F.Plan(type_, modes, 1, tol, isign, dtype, upsampfac=2.0, gpu_device_id=local)
            full_pts = perfdata.points(3, M, rt, args.dist, nf)
            fx, fy, fz = (torch.from_numpy(p).to(dev) for p in full_pts)
            plan.setpts(fz, fy, fx)
            shp = (M,) if type_ == 1 else modes
            dd = torch.view_as_complex(torch.randn(shp + (2,), dtype=rdt, device=dev))
            oo = plan.execute(dd)
            ms_w = timed(lambda i: plan.execute(dd, oo), max(5, args.steps // 2), 2)
            replicas = {"value": world * M / (ms_w * 1e-3), "unit": "points/s",
                        "ms_per_step": ms_w, "scaling": "weak",
                        "workload": f"{world} independent transforms of M points"}
            plan.destroy()
        if rank == 0:
            achieved = abytes / (kernel_ms * 1e-3) / 1e9
            line = {"metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                    "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": prec, "data": "synthetic", "config": config, "clocks": clocks,
                    "e2e": {"value": M / (e2e_ms * 1e-3), "unit": "points/s",
                            "h2d_bytes_per_step": int(world * h_in[0].numel() * cbytes),
                            "d2h_bytes_per_step": int(world * h_out.numel() * cbytes
                                                      if type_ == 1 else M * cbytes),
                            "ms_per_step": e2e_ms,
                            "api": "b200_slab[f]_execute: pinned host -> device copy of the "
                                   "rank's input, sharded execute, device -> host of its output"},
                    # this library's own kernels launched inside the timed region (cuFFT's and
                    # memsets not counted), this rank
                    "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches,
                    "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved,
                                 "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                 "traffic": None, "peak_source": peak_kind,
                                 "algorithmic_bytes": abytes, "kernel_ms": kernel_ms,
                                 "note": "per rank: its routed points and its window of the grid"},
                    "cpu_baseline": None, "stages_ms": stage_avg, "setpts_ms": setpts_ms,
                    "setpts_wall_ms": setpts_wall_ms,
                    "value_with_setpts": M / ((ms_per_step + setpts_ms) * 1e-3),
                    "accuracy": acc,
                    "plan": {"ns": info["ns"], "nf": nf, "mode": info["mode"],
                             "window_planes": info["win_n"], "M_local_rank0": info["M_local"],
                             "points": "resident on the owning GPU" if resident else
                                       "arbitrary share per rank (replicated-window mode)"}}
            line.update(extras)
            if replicas:
                line["replicas"] = replicas
    else:
        # ============================================================== one plan per GPU
        ntr_l = ntr
        if batched:
            from finufft_b200.parallel import split_transforms
            lo_v, hi_v = split_transforms(ntr, world)[rank]
            ntr_l = hi_v - lo_v
        plan = F.Plan(type_, modes, max(ntr_l, 1), tol, isign, dtype, upsampfac=2.0,
                      gpu_device_id=local)
        info = plan.info()
        nf = info["nf"]
        pts_h = perfdata.points(dim, M, rt, args.dist, nf)
        pts = [torch.from_numpy(p).to(dev) for p in pts_h]
        plan.enable_profiling(True)
        plan.setpts(*pts[::-1])  # untimed: first-call module loading, pool growth
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        plan.setpts(*pts[::-1])
        torch.cuda.synchronize()
        setpts_wall_ms = (time.perf_counter() - t0) * 1e3
        setpts_ms = plan.stage_ms()["setpts"]
        info = plan.info()
        nb = (max(ntr_l, 1),) if max(ntr_l, 1) > 1 else ()
        in_shape = nb + ((M,) if type_ == 1 else tuple(modes))
        out_shape = nb + (tuple(modes) if type_ == 1 else (M,))
        if ntr == 1:
            n_in = M if type_ == 1 else nmodes
            h_in = [torch.from_numpy(perfdata.strengths(n_in, dtype, s).reshape(in_shape))
                    .pin_memory() for s in (("C", "FK") if type_ == 1 else ("FK", "C"))]
            data = [h.to(dev) for h in h_in]
        else:  # many vectors: generated on the device (10 GB at C4), uniform in (-1,1)
            gen = torch.Generator(device=dev)
            gen.manual_seed(99 + rank)
            data = [torch.view_as_complex(
                torch.rand(in_shape + (2,), dtype=rdt, device=dev, generator=gen) * 2 - 1)]
            data.append(data[0])
            h_in = None
        out = plan.execute(data[0])
        torch.cuda.synchronize()
        if rank == 0:
            sampler.start()
        l0 = plan.launch_count()
        ms_per_step = timed(lambda i: plan.execute(data[i & 1], out), args.steps, args.warmup)
        launches = (plan.launch_count() - l0) // (args.steps + args.warmup)
        clocks = sampler.stop() if rank == 0 else None
        total_vectors = ntr if (batched or world == 1) else world * ntr
        value = M * total_vectors / (ms_per_step * 1e-3)
        stage = {"spreadinterp": [], "fft": [], "deconv": []}
        for i in range(max(3, min(args.steps, 10))):
            plan.execute(data[i & 1], out)
            s = plan.stage_ms()
            for k in stage:
                stage[k].append(s[k])
        stage_avg = {k: float(np.mean(v)) for k, v in stage.items()}
        nbatch = -(-max(ntr_l, 1) // info["batch"])  # stage times are those of the last batch
        plan.execute(data[0], out)
        first_in = data[0][0] if nb else data[0]
        first_out = out[0] if nb else out
        if type_ == 1:
            acc = direct_sum_check(torch, 1, modes, isign, pts, first_in, first_out)
        else:
            acc = direct_sum_check(torch, 2, modes, isign, pts, first_out, first_in)

        # other transform type on the same points (the metric is quoted for type 1 and 2)
        other = None
        if world == 1 and ntr == 1 and not args.no_extras and dim >= 2:
            t2 = 3 - type_
            p2 = F.Plan(t2, modes, 1, tol, isign, dtype, upsampfac=2.0, gpu_device_id=local)
            p2.enable_profiling(True)
            p2.setpts(*pts[::-1])
            d2 = [out, out]
            o2 = p2.execute(out)
            ms2 = timed(lambda i: p2.execute(d2[i & 1], o2), max(5, args.steps // 2), 3)
            s2 = p2.stage_ms()
            other = {"type": t2, "value": M / (ms2 * 1e-3), "unit": "points/s",
                     "ms_per_step": ms2, "stages_ms": {k: s2[k] for k in stage}}
            p2.destroy()
            del o2

        # ---------------------------------------------------------------- e2e: host-pointer ABI
        hplan = F.HostPlan(type_, modes, max(ntr_l, 1), tol, isign, dtype, upsampfac=2.0,
                           allow_eps_too_small=1)
        hplan.setpts(*pts_h[::-1])
        if h_in is None:
            h_in = [torch.empty(in_shape, dtype=cdt).pin_memory()]
            h_in[0].copy_(data[0].cpu())
            h_in.append(h_in[0])
        h_out = torch.empty(out_shape, dtype=cdt).pin_memory()
        h_np = [h.numpy() for h in h_in]
        h_out_np = h_out.numpy()
        e2e_steps = args.steps if ntr == 1 else max(2, min(args.steps, 5))
        e2e_ms = timed_wall(lambda i: hplan.execute(h_np[i & 1], out=h_out_np), e2e_steps, 1)
        e2e_val = M * total_vectors / (e2e_ms * 1e-3)
        hplan.destroy()

        if rank == 0:
            cells = int(np.prod(nf))
            abytes = algorithmic_bytes(dim, M, cells, rbytes)
            # stage times are per batch of info["batch"] vectors: one launch per vector
            nlast = max(ntr_l, 1) - (nbatch - 1) * info["batch"]
            kernel_ms = stage_avg["spreadinterp"] / nlast
            sweep_on = os.environ.get("B200_NUFFT_SWEEP", "1") != "0"
            tf = "true" if type_ == 1 else "false"
            if sweep_on and dim == 3 and rbytes == 4:
                kernel_name = f"k_sweep3<{info['ns']},{tf}>"
            elif sweep_on and dim == 2:
                kernel_name = f"k_sweep2<{'float' if rbytes == 4 else 'double'},{info['ns']},{tf}>"
            else:
                kernel_name = "k_spread" if type_ == 1 else "k_interp"
            achieved = abytes / (kernel_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": kernel_name,
                        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": ncu_traffic(args.workload, args.dist),
                        "traffic_source": "profiles/traffic.json (one ncu --set full capture, "
                                          "not measured in this run)",
                        "peak_source": peak_kind, "algorithmic_bytes": abytes,
                        "kernel_ms": kernel_ms}
            # secondary roof (SURVEY.md 8(d) "which roof binds"): ns^d complex cell updates per
            # point against the packed-FFMA2 rate measured on this part (tools/micro/micro1.cu,
            # profiles/r1_micro_ffma2_red.txt: 17.6e12 cell updates/s)
            if rbytes == 4:
                ncell = float(info["ns"]) ** dim * M
                roofline["fp32"] = {"achieved_tcell_s": ncell / (kernel_ms * 1e-3) / 1e12,
                                    "peak_tcell_s": 17.6,
                                    "frac": ncell / (kernel_ms * 1e-3) / 17.6e12,
                                    "floor_ms": ncell / 17.6e12 * 1e3}
            cpu = None
            if not args.no_cpu and world == 1:
                cmodes, cM, what = cpu_sample_shape(modes, M, ntr)
                try:
                    sec, _, nthr, kind = cpu_run(type_, cmodes, cM, tol, dtype, 1, 0, args.dist)
                    cpu = {"value": cM / sec, "unit": "points/s", "cores": nthr, "kind": kind,
                           "sample": f"{what}: modes {'x'.join(map(str, cmodes))}, M={cM:.3g}, "
                                     "1 execute"}
                except Exception as exc:  # the checker failing must not lose the GPU number
                    cpu = {"value": None, "unit": "points/s", "cores": 0, "kind": "port",
                           "sample": f"failed: {exc}"}
            line = {"metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                    "higher_is_better": True, "scaling": "strong" if batched else "weak",
                    "vs_baseline": None, "dtype": prec, "data": "synthetic", "config": config,
                    "clocks": clocks,
                    "e2e": {"value": e2e_val, "unit": "points/s",
                            "h2d_bytes_per_step": int(np.prod(in_shape)) * cbytes * (1 if batched or world == 1 else world),
                            "d2h_bytes_per_step": int(np.prod(out_shape)) * cbytes * (1 if batched or world == 1 else world),
                            "ms_per_step": e2e_ms, "steps": e2e_steps,
                            "api": "finufft[f]_execute (host pointers, pinned buffers)"},
                    # this library's own kernels launched inside the timed region (cuFFT's and
                    # memsets not counted), this rank
                    "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches,
                    "roofline": roofline, "cpu_baseline": cpu,
                    "stages_ms": stage_avg, "setpts_ms": setpts_ms,
                    "setpts_wall_ms": setpts_wall_ms,
                    "value_with_setpts": M * total_vectors / ((ms_per_step + setpts_ms) * 1e-3),
                    "accuracy": acc,
                    "plan": {"ns": info["ns"], "nf": nf, "nsub": info["nsub"],
                             "batch": info["batch"], "vectors_this_rank": max(ntr_l, 1)}}
            if other:
                line[f"value_t{other['type']}"] = other["value"]
                line[f"type{other['type']}"] = other
        plan.destroy()

    if rank == 0:
        # sanity bound on the timed plan's own output (not the parity gate: see tests/)
        bound = 1e-3 if rbytes == 4 else 1e-7
        line["accuracy"]["bound"] = bound
        if not (line["accuracy"]["relerr"] <= bound):
            line["accuracy"]["failed"] = True
            print(json.dumps(line), flush=True)
            raise SystemExit(f"accuracy check failed: {line['accuracy']}")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist_.barrier()
        dist_.destroy_process_group()


if __name__ == "__main__":
    main()
