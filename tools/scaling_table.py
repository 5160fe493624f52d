"""Scaling tables of DESIGN.md section 6 from profiles/r2_bench_*_<N>gpu.json and the 1-GPU lines
profiles/<tag>_bench_*.json.   python tools/scaling_table.py [1gpu-tag]"""
import glob, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
one = sys.argv[1] if len(sys.argv) > 1 else "r2z"


def load(p):
    try:
        return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception:
        return None


b = load(os.path.join(P, f"{one}_bench_c3_t1.json"))
print("**C3 type 1 (3D f32, 256³ modes, M = 1e8, tol 1e-6): ONE transform over N GPUs** "
      "(`profiles/r2_bench_c3_t1_<N>gpu.json`)\n")
print("| GPUs | ms / step, inputs resident on the owning GPU | NU pts/s | x vs 1 GPU | arbitrary points, routed inside the step: ms (x) | + gather of the modes onto every rank: ms | e2e (pinned host in / out): ms, pts/s | setpts ms (resident / routed) | N replicas (weak): pts/s |")
print("|---|---|---|---|---|---|---|---|---|")
print(f"| 1 | {b['ms_per_step']:.2f} | {b['value']:.3g} | 1.00 | - | - | {b['e2e']['ms_per_step']:.2f}, {b['e2e']['value']:.3g} | {b['setpts_ms']:.2f} | {b['value']:.3g} |")
stages = {}
for n in (2, 4, 8):
    d = load(os.path.join(P, f"r2_bench_c3_t1_{n}gpu.json"))
    if not d:
        continue
    a, g, r = d.get("arbitrary_points", {}), d.get("with_mode_gather", {}), d.get("replicas", {})
    print(f"| {n} | {d['ms_per_step']:.2f} | {d['value']:.3g} | {d['value'] / b['value']:.2f} | "
          f"{a.get('ms_per_step', float('nan')):.2f} ({b['ms_per_step'] / a.get('ms_per_step', float('nan')):.2f}x) | "
          f"{g.get('ms_per_step', float('nan')):.2f} | {d['e2e']['ms_per_step']:.2f}, {d['e2e']['value']:.3g} | "
          f"{d['setpts_ms']:.2f} / {a.get('setpts_ms', float('nan')):.2f} | {r.get('value', float('nan')):.3g} |")
    stages[n] = (d["stages_ms"], a.get("route_values_ms"))
print("\nStages of one step, ms, max over ranks (CUDA events on each rank's stream):\n")
keys = ["spreadinterp", "ghost", "fft2d", "pack", "transpose", "fft1d", "deconv"]
print("| GPUs | " + " | ".join(keys) + " | route_values (arbitrary points only) |")
print("|---|" + "---|" * (len(keys) + 1))
for n, (st, rv) in stages.items():
    print(f"| {n} | " + " | ".join(f"{st[k]:.3f}" for k in keys) + f" | {rv:.3f} |")
print("\n(`ghost` = barrier + the two adds from the neighbours' windows; `pack` = crop + stores into the "
      "owners' pencils over NVLink; `transpose` = the barrier that closes it.)\n")
print("**Other workloads** (`profiles/r2_bench_*`)\n")
print("| workload | GPUs | ms / step | NU pts/s | x vs 1 GPU | e2e pts/s |")
print("|---|---|---|---|---|---|")
for name, files, base in (
        ("C3 type 2, arbitrary points routed inside the step", [(8, "r2_bench_c3_t2_8gpu_routed.json")], "c3_t2"),
        ("C3 type 1, clustered points (replicated window)", [(2, "r2_bench_c3_t1_cluster_2gpu.json"), (8, "r2_bench_c3_t1_cluster_8gpu.json")], "c3_t1_cluster"),
        ("C4 type 1 (2D f64, ntransf = 64 split over the GPUs)", [(2, "r2_bench_c4_t1_2gpu.json"), (4, "r2_bench_c4_t1_4gpu.json"), (8, "r2_bench_c4_t1_8gpu.json")], "c4_t1")):
    bb = load(os.path.join(P, f"{one}_bench_{base}.json"))
    print(f"| {name} | 1 | {bb['ms_per_step']:.2f} | {bb['value']:.3g} | 1.00 | {bb['e2e']['value']:.3g} |")
    for n, f in files:
        d = load(os.path.join(P, f))
        if d:
            print(f"| | {n} | {d['ms_per_step']:.2f} | {d['value']:.3g} | {d['value'] / bb['value']:.2f} | {d['e2e']['value']:.3g} |")
