"""Static evidence for every kernel of the built library: registers, static shared memory, stack
(spills), SASS length and the counts of the mnemonics that say what the kernel does (packed /
scalar FMA, reductions, atomics, shared-memory and global traffic, shuffles, barriers).
Reads finufft_b200/build/*.o with cuobjdump (no GPU needed).

    python tools/sass_index.py r2zz          # -> profiles/r2zz_sass_index.txt
                                             #    profiles/r2zz_sass_<kernel>.txt for the hot kernels
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2zz"
B, P = os.path.join(ROOT, "finufft_b200", "build"), os.path.join(ROOT, "profiles")
KEYS = ["FFMA2", "FMUL2", "FFMA", "DFMA", "RED", "ATOM", "LDG", "STG", "LDS", "STS", "SHFL",
        "MATCH", "BAR", "LDGSTS", "BRA"]
# full listings: the kernels of one C3 / C2 setpts + execute
LIST = {
    "k_sweep3_spread_ns7": "_ZN4b2008k_sweep3ILi7ELb1EEEvNS_9SweepArgsIXT_EEE",
    "k_sweep3_interp_ns7": "_ZN4b2008k_sweep3ILi7ELb0EEEvNS_9SweepArgsIXT_EEE",
    "k_sweep2_spread_f32_ns6": "_ZN4b2008k_sweep2IfLi6ELb1EEEvNS_10Sweep2ArgsIT_XT0_EEE",
    "k_sweep2_interp_f32_ns6": "_ZN4b2008k_sweep2IfLi6ELb0EEEvNS_10Sweep2ArgsIT_XT0_EEE",
    "k_sweep2_spread_f64_ns10": "_ZN4b2008k_sweep2IdLi10ELb1EEEvNS_10Sweep2ArgsIT_XT0_EEE",
    "k_bin_hist_f32_3d": "_ZN4b20010k_bin_histIfLi3EEEvPKT_S3_S3_jNS_8GridGeomIS1_EEPj",
    "k_part_f32_3d_raw": "_ZN4b2006k_partIfLi3ELb1EEEvPKT_S3_S3_PKNS_7Packed4IS1_EEjNS_8GridGeomIS1_EEiiPjPS5_",
    "k_seg_sort_f32_3d": "_ZN4b20010k_seg_sortIfLi3ELi1EEEvPKNS_7Packed4IT_EEPKjjNS_8GridGeomIS2_EEiiiPS2_SA_SA_Pj",
}


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout
    return dict(zip(names, out.splitlines()))


rows, listings = [], {}
for obj in sorted(glob.glob(os.path.join(B, "*.o"))):
    res = subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True).stdout
    usage, cur = {}, None
    for ln in res.splitlines():
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in ln:
            usage[cur] = dict(kv.split(":") for kv in ln.split() if ":" in kv and "[" not in kv)
            cur = None
    if not usage:
        continue
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, counts, length, text = None, {}, {}, collections.defaultdict(list)
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur], length[cur] = collections.Counter(), 0
            continue
        s = ln.strip()
        if cur and s.startswith("/*") and "*/" in s[:12]:
            body = s.split("*/", 1)[1].split("/* 0x")[0].strip().rstrip(";").strip()
            if not body:
                continue
            ops = [w for w in body.split() if not w.startswith("@")]
            if not ops:
                continue
            op = ops[0].split(".")[0]
            length[cur] += 1
            counts[cur][op] += 1
            text[cur].append(s.split("/* 0x")[0].rstrip())
    for f in usage:
        rows.append((os.path.basename(obj), f, usage[f], length.get(f, 0), counts.get(f, {})))
    for name, sym in LIST.items():
        if sym in text:
            listings[name] = text[sym]

dm = demangle([r[1] for r in rows])
rows.sort(key=lambda r: (r[0], dm[r[1]]))
hdr = f"{'kernel':<78} {'regs':>4} {'smem':>6} {'stack':>5} {'SASS':>5} " + " ".join(f"{k:>6}" for k in KEYS)
lines = [f"# {tag}: every __global__ function of finufft_b200/libfinufft_b200.so (sm_100a), from cuobjdump -res-usage / -sass",
         "# of the objects of the final build.  smem = STATIC shared memory (the sweep / partition kernels add dynamic",
         "# shared memory at launch), stack > 0 = spills.  FFMA2 / FMUL2 = packed FP32 pairs, RED = fire-and-forget",
         "# reductions into the fine grid, ATOM = atomics that return a value (cursors), MATCH = warp match_any.",
         hdr]
last = None
for obj, f, u, n, c in rows:
    if obj != last:
        lines.append(f"## {obj}")
        last = obj
    name = dm[f].replace("(anonymous namespace)::", "").replace("b200::", "").replace("void ", "")
    name = re.sub(r"\(.*", "", name)
    lines.append(f"{name[:78]:<78} {u.get('REG', '?'):>4} {u.get('SHARED', '?'):>6} {u.get('STACK', '?'):>5} {n:>5} "
                 + " ".join(f"{c.get(k, 0):>6}" for k in KEYS))
open(os.path.join(P, f"{tag}_sass_index.txt"), "w").write("\n".join(lines) + "\n")
for name, t in listings.items():
    open(os.path.join(P, f"{tag}_sass_{name}.txt"), "w").write("\n".join(t) + "\n")
print(f"{len(rows)} kernels -> profiles/{tag}_sass_index.txt; listings: {', '.join(listings)}")
