#!/bin/bash
# One GPU-box visit: parity tests, bench lines for every single-GPU workload, ncu launch lists and
# `--set full` captures of the dominant kernels.  Everything lands in gpurun_out/.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh [tag]'
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/${tag}_smi.txt 2>&1
nproc >> $out/${tag}_smi.txt
grep -m1 "model name" /proc/cpuinfo >> $out/${tag}_smi.txt
timeout 1500 python -m pytest tests -m gpu -q -s > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
grep -E "passed|failed|FAILED|gpu-vs" $out/${tag}_pytest_gpu.log | tail -40
timeout 600 python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_c3_t1.json 2> $out/${tag}_bench_c3_t1.err
tail -c 3000 $out/${tag}_bench_c3_t1.json; tail -3 $out/${tag}_bench_c3_t1.err
for w in c3_t2 c2_t2 c2_t1 c1_t1; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err
done
timeout 600 python bench.py --workload c4_t1 --steps 5 --warmup 3 --no-cpu > $out/${tag}_bench_c4_t1.json 2> $out/${tag}_bench_c4_t1.err
for w in c3_t1 c3_t2 c2_t2; do
  timeout 400 python bench.py --workload $w --dist cluster --steps 10 --warmup 3 --no-cpu --no-extras > $out/${tag}_bench_${w}_cluster.json 2> $out/${tag}_bench_${w}_cluster.err
done
timeout 300 python tools/bench_type3.py > $out/${tag}_bench_c5_t3.txt 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
tail -c 1200 $out/${tag}_bench_reference.json
# launch list of the bench command (per-launch times are cold-cache; the SHARE is what counts)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $out/${tag}_launches_c3_t1.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > $out/${tag}_ncu_bench.log 2>&1
# launch list + DRAM bytes of one setpts + execute
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
  --log-file $out/${tag}_launches_setpts.csv python tools/prof_run.py --workload c3_t1 --reps 1 > $out/${tag}_ncu_setpts.log 2>&1
# full capture of the dominant kernels (one launch each)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep3 -c 2 -f \
  -o $out/${tag}_sweep_spread python tools/prof_run.py --workload c3_t1 --reps 2 > $out/${tag}_ncu_full_spread.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep3 -c 1 -f \
  -o $out/${tag}_sweep_interp python tools/prof_run.py --workload c3_t2 --reps 1 > $out/${tag}_ncu_full_interp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep2 -c 1 -f \
  -o $out/${tag}_sweep2_interp python tools/prof_run.py --workload c2_t2 --reps 1 > $out/${tag}_ncu_full_interp2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep2 -c 1 -f \
  -o $out/${tag}_sweep2_spread python tools/prof_run.py --workload c2_t1 --reps 1 > $out/${tag}_ncu_full_spread2.log 2>&1
# setpts kernels: one full capture each (first launch of k_part = raw pass, second = record pass)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_part|k_seg_sort|k_bin_hist" -c 4 -f \
  -o $out/${tag}_setpts python tools/prof_run.py --workload c3_t1 --reps 1 > $out/${tag}_ncu_full_setpts.log 2>&1
# the sharded plan at world = 1: launch list of the slab kernels
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv \
  --log-file $out/${tag}_launches_slab.csv python tools/prof_run_sharded.py > $out/${tag}_ncu_slab.log 2>&1
tail -3 $out/${tag}_ncu_slab.log
ls -la $out | tail -40
