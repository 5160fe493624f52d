#!/bin/bash
out=gpurun_out; mkdir -p $out
tag=${1:-n2d}
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_sweep2 -c 1 -f \
  -o $out/${tag}_interp python tools/prof_run.py --workload c2_t2 --reps 1 > $out/${tag}_interp.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_sweep2 -c 1 -f \
  -o $out/${tag}_spread python tools/prof_run.py --workload c2_t1 --reps 1 > $out/${tag}_spread.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_refine|k_bin_place|k_bin_count' -c 3 -f \
  -o $out/${tag}_setpts python tools/prof_run.py --workload c3_t1 --reps 1 > $out/${tag}_setpts.log 2>&1
ls -la $out | tail -5
