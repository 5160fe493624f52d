"""Copy one gpu_round.sh visit from gpurun_out/ into profiles/ (tracked) and derive the summaries:
ncu raw/source digests, launch shares of one C3 type-1 execute, traffic.json, SASS listings.
    python tools/collect_profiles.py r1c
"""
import collections
import csv
import glob
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2z"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

for f in glob.glob(os.path.join(G, f"{tag}_bench_*.json")) + glob.glob(os.path.join(G, f"{tag}_bench_*.txt")):
    shutil.copy(f, P)
for f in (f"{tag}_pytest_gpu.log", f"{tag}_launches_c3_t1.csv", f"{tag}_launches_setpts.csv", f"{tag}_launches_slab.csv",
          f"{tag}_smi.txt"):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), P)

# ncu digests
for k in ("sweep_spread", "sweep_interp", "sweep2_spread", "sweep2_interp"):
    rep = os.path.join(G, f"{tag}_{k}.ncu-rep")
    if os.path.exists(rep):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, "1e8"],
                             capture_output=True, text=True).stdout
        open(os.path.join(P, f"{tag}_{k}_ncu.txt"), "w").write(out.replace(G + "/", "gpurun_out/"))


rep = os.path.join(G, f"{tag}_setpts.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_digest.py"), rep],
                         capture_output=True, text=True).stdout
    open(os.path.join(P, f"{tag}_setpts_ncu.txt"), "w").write(
        f"# ncu --set full of the setpts kernels at C3 (gpurun_out/{tag}_setpts.ncu-rep)\n" + out)


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(out)))
    d = dict(zip(rr[0], zip(rr[1], rr[2])))

    def b(k):
        u, v = d[k]
        return float(v.replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
    return b("dram__bytes_read.sum"), b("dram__bytes_write.sum"), d["Kernel Name"][1]


traffic = {}
for key, k in (("c3_t1:uniform", "sweep_spread"), ("c3_t2:uniform", "sweep_interp"),
               ("c2_t1:uniform", "sweep2_spread"), ("c2_t2:uniform", "sweep2_interp")):
    rep = os.path.join(G, f"{tag}_{k}.ncu-rep")
    if os.path.exists(rep):
        r, w, name = raw(rep)
        traffic[key] = {"dram_bytes": r + w, "read": r, "write": w, "kernel": name,
                        "source": f"profiles/{tag}_{k}_ncu.txt"}
if traffic:
    json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)

# launch shares of one device-resident execute
lc = os.path.join(G, f"{tag}_launches_c3_t1.csv")
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if len(r) > 5]
    H = rows[0]
    ik, iv = H.index("Kernel Name"), H.index("Metric Value")
    ms = collections.defaultdict(list)
    for r in rows[1:]:
        ms[r[ik].split("(")[0].replace("void ", "").replace("b200::", "")].append(float(r[iv].replace(",", "")) / 1e6)
    sw = ms.get("k_sweep3<7, 1>", [])
    full = [x for x in sw if x > 0.75 * max(sw)] if sw else []
    grp = [x for x in sw if x <= 0.75 * max(sw)] if sw else []
    g2m = ms.get("k_grid_to_modes<float, 3>", [])
    nexec = max(len(g2m), 1)
    fft = sum(sum(v) for k, v in ms.items() if "_fft" in k) / nexec
    sp = sum(full) / max(len(full), 1)
    dc = sum(g2m) / nexec
    tot = sp + fft + dc
    bench = json.loads(open(os.path.join(G, f"{tag}_bench_c3_t1.json")).read().strip().splitlines()[-1])
    st = bench["stages_ms"]
    setp = {k: sum(v) / len(v) for k, v in ms.items()
            if k.startswith(("k_bin_hist", "k_part", "k_seg_sort", "k_bin_count", "k_bin_place", "k_refine"))}
    txt = f"""# {tag}: ncu --metrics gpu__time_duration.sum --clock-control none, `python bench.py --steps 2 --warmup 3 --no-cpu`
# (profiles/{tag}_launches_c3_t1.csv; per-launch times are cold-cache/serialised: the SHARE of the step counts)
# one device-resident execute of C3 type 1 = memset(fw) + k_sweep3<7,1> + cuFFT kernels (pruned 3D) + k_grid_to_modes
kernel                     per-execute ms   share of execute kernels
k_sweep3<7,1> (spread)     {sp:.3f}            {100 * sp / tot:.1f} %
cuFFT (all kernels)        {fft:.3f}            {100 * fft / tot:.1f} %
k_grid_to_modes<float,3>   {dc:.3f}            {100 * dc / tot:.1f} %
# bench.py stages_ms (CUDA events, same run family): spreadinterp {st['spreadinterp']:.3f} (incl. 0.17 ms memset), fft {st['fft']:.3f}, deconv {st['deconv']:.3f}
# the e2e leg (finufft_execute, host pointers) spreads in 4 point groups: {len(grp)} k_sweep3 launches of {min(grp or [0]):.2f}..{max(grp or [0]):.2f} ms (and the type-2 side leg's interp launches)
# setpts (once per point set): """ + ", ".join(f"{k} {v:.3f}" for k, v in setp.items()) + "\n"
    open(os.path.join(P, f"{tag}_launch_shares.txt"), "w").write(txt)
    print(txt)

# SASS listings of the built objects
objs = {"k_sweep3_spread_ns7": ("sweep3d.o", "_ZN4b2008k_sweep3ILi7ELb1EEEvNS_9SweepArgsIXT_EEE"),
        "k_sweep2_spread_f32_ns6": ("sweep2d_f32.o", "_ZN4b2008k_sweep2IfLi6ELb1EEEvNS_10Sweep2ArgsIT_XT0_EEE"),
        "k_sweep2_interp_f32_ns6": ("sweep2d_f32.o", "_ZN4b2008k_sweep2IfLi6ELb0EEEvNS_10Sweep2ArgsIT_XT0_EEE")}
for name, (obj, sym) in objs.items():
    o = os.path.join(ROOT, "finufft_b200", "build", obj)
    if not os.path.exists(o):
        continue
    out = subprocess.run(["cuobjdump", "-sass", o, "-fun", sym], capture_output=True, text=True).stdout
    lines = [ln.strip().split("/* 0x")[0].rstrip() for ln in out.splitlines()
             if ln.strip().startswith("/*") and "*/" in ln.strip()[:12]]
    open(os.path.join(P, f"{tag}_sass_{name}.txt"), "w").write("\n".join(lines) + "\n")
print("profiles/ updated for", tag)
