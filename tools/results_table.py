"""Markdown table of the bench lines in profiles/<tag>_bench_*.json (for DESIGN.md section 5)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2z"
rows = [("c3_t1", "C3 type 1 (headline)"), ("c3_t2", "C3 type 2"), ("c3_t1_cluster", "C3 type 1, clustered"),
        ("c3_t2_cluster", "C3 type 2, clustered"), ("c2_t2", "C2 type 2"), ("c2_t1", "2D 2048², M=1e8, type 1"),
        ("c4_t1", "C4 type 1, ntransf = 64 (ms per 64 vectors; kernel = per vector)"), ("c2_t2_cluster", "C2 type 2, clustered"), ("c1_t1", "C1 type 1 (1D, double)")]
print("| workload | NU pts/s (device-resident) | ms / execute | spread or interp ms | FFT ms | deconv / amplify ms | "
      "kernel GB/s (algorithmic) = frac of HBM | e2e pts/s (host pointers) | setpts ms | rel. error of the timed output vs f64 direct sum |")
print("|---|---|---|---|---|---|---|---|---|---|")
for key, name in rows:
    p = os.path.join(ROOT, "profiles", f"{tag}_bench_{key}.json")
    if not os.path.exists(p):
        continue
    d = json.loads(open(p).read().strip().splitlines()[-1])
    st, rf = d["stages_ms"], d["roofline"]
    acc = d.get("accuracy", {}).get("relerr")
    print(f"| {name} | {d['value']:.3g} | {d['ms_per_step']:.2f} | {st['spreadinterp']:.2f} | {st['fft']:.2f} | "
          f"{st['deconv']:.3f} | {rf['achieved']:.0f} = {rf['frac']:.3f} | {d['e2e']['value']:.3g} | {d['setpts_ms']:.2f} | "
          + (f"{acc:.1e} |" if acc is not None else "- |"))
p = os.path.join(ROOT, "profiles", f"{tag}_bench_reference.json")
if os.path.exists(p):
    d = json.loads(open(p).read().strip().splitlines()[-1])
    cb = d["cpu_baseline"]
    print(f"\nCPU arm (`bench.py --impl reference`, kind = {cb['kind']}, {cb['cores']} cores): {d['value']:.3g} pts/s "
          f"({cb['sample']}).")
p = os.path.join(ROOT, "profiles", f"{tag}_bench_c5_t3.txt")
if os.path.exists(p):
    d = json.loads(open(p).read().strip().splitlines()[-1])
    print(f"C5 (3D type 3, M=N=1e7): execute {d['execute_ms']:.1f} ms = {d['points_per_s']:.3g} pts/s, setpts {d['setpts_ms']:.1f} ms.")
