"""Raw-metric digest of every launch in an .ncu-rep (read here, no GPU): time, DRAM bytes, occupancy limits, issue / shared-memory utilisation, top stall reasons.
usage: python tools/ncu_digest.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv,sys,io,subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
H,U=rows[0],rows[1]
want=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","launch__registers_per_thread","launch__grid_size","launch__block_size","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","sm__throughput.avg.pct_of_peak_sustained_elapsed","launch__occupancy_limit_shared_mem","launch__occupancy_limit_registers","lts__t_sectors_op_atom.sum","lts__t_sectors_op_red.sum","smsp__inst_executed.sum","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]
for V in rows[2:]:
    for h,u,v in zip(H,U,V):
        if h in want: print(h[:75],'['+u+']',v[:100])
    st=[(float(v),h) for h,u,v in zip(H,U,V) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    st.sort(reverse=True)
    print("top stalls:", [(h.split("stalled_")[1].split("_per")[0],round(x,2)) for x,h in st[:6]])
    print()
