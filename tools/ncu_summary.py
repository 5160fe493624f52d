"""Summarise an .ncu-rep (read here, no GPU): key raw metrics + instruction/stall breakdown by SASS region.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [points]  > profiles/x.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
npts = float(sys.argv[2]) if len(sys.argv) > 2 else 1e8
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U, V = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_red.sum",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
print(f"# {rep}")
for h, u, v in zip(H, U, V):
    if h in want or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
H = rows[1]; data = rows[2:]
ia, isamp, isrc = H.index("Instructions Executed"), H.index("# Samples"), H.index("Source")
for k, r in enumerate(data):  # a report with several launches repeats the table: keep the first
    if len(r) <= max(ia, isamp, isrc) or not r[ia].strip().isdigit():
        data = data[:k]; break
tot = sum(int(r[ia]) for r in data); tots = sum(int(r[isamp]) for r in data)
print(f"\n# warp-instructions per point = {tot / npts:.2f}; stall samples = {tots}")
print("# SASS regions (consecutive lines with equal execution count): lines, instr/pt, share of samples, top opcodes")
prev, grp = None, []
def flush(grp):
    if not grp: return
    n = sum(int(r[ia]) for _, r in grp); s = sum(int(r[isamp]) for _, r in grp)
    if n / npts < 0.3 and s / max(tots, 1) < 0.01: return
    ops = {}
    for _, r in grp:
        t = r[isrc].split(); op = t[0] if not t[0].startswith("@") else t[1]
        ops[op] = ops.get(op, 0) + 1
    top = ", ".join(f"{k}x{v}" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:5])
    print(f"{grp[0][0]:5d}-{grp[-1][0]:5d}  {n / npts:7.2f}/pt  {100 * s / max(tots, 1):5.1f}%  {top}")
for i, r in enumerate(data):
    n = int(r[ia])
    if prev is not None and abs(n - prev) > 0.05 * max(n, prev, 1): flush(grp); grp = []
    grp.append((i, r)); prev = n
flush(grp)
