"""Copies the multi-GPU bench lines of gpurun_out/<tag>_bench_<name>_g<N>.json into profiles/ as
r2_bench_<name>_<N>gpu.json and prints the scaling tables of DESIGN.md section 6.
    python tools/scaling_table.py 2:r2i 4:r2j 8:r2k [1gpu-tag]"""
import json, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
runs = dict(a.split(":") for a in sys.argv[1:] if ":" in a)
one = [a for a in sys.argv[1:] if ":" not in a]
one = one[0] if one else "r2g"


def load(p):
    try:
        return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception:
        return None


base = {k: load(os.path.join(P, f"{one}_bench_{k}.json")) for k in ("c3_t1", "c3_t2", "c3_t1_cluster", "c4_t1")}
for name, title in (("c3_t1", "C3 type 1, one transform sharded (strong)"), ("c3_t2", "C3 type 2, sharded"),
                    ("c3_t1_cluster", "C3 type 1, clustered points (replicated window)"),
                    ("c4_t1", "C4 type 1, 64 vectors split over the GPUs")):
    b = base.get(name)
    print(f"\n**{title}**\n")
    print("| GPUs | ms / step | NU pts/s | x vs 1 GPU | e2e ms | e2e pts/s | pre-partitioned ms (x) | with mode gather ms | stages ms (max over ranks) |")
    print("|---|---|---|---|---|---|---|---|---|")
    if b:
        print(f"| 1 | {b['ms_per_step']:.2f} | {b['value']:.3g} | 1.00 | {b['e2e']['ms_per_step']:.2f} | {b['e2e']['value']:.3g} | - | - | "
              + ", ".join(f"{k} {v:.2f}" for k, v in b['stages_ms'].items()) + " |")
    for n in sorted(runs, key=int):
        src = os.path.join(G, f"{runs[n]}_bench_{name}_g{n}.json")
        d = load(src)
        if not d:
            continue
        shutil.copy(src, os.path.join(P, f"r2_bench_{name}_{n}gpu.json"))
        x = d['value'] / b['value'] if b else float('nan')
        pp = d.get('pre_partitioned_points') or {}
        gm = d.get('with_mode_gather') or {}
        ppx = f"{pp['ms_per_step']:.2f} ({b['ms_per_step'] / pp['ms_per_step']:.2f}x)" if 'ms_per_step' in pp and b else "-"
        st = ", ".join(f"{k} {v:.2f}" for k, v in d['stages_ms'].items() if k != 'execute')
        print(f"| {n} | {d['ms_per_step']:.2f} | {d['value']:.3g} | {x:.2f} | {d['e2e']['ms_per_step']:.2f} | {d['e2e']['value']:.3g} | "
              f"{ppx} | {gm.get('ms_per_step', float('nan')):.2f} | {st} |")
