#!/bin/bash
# One GPU-box visit: parity tests, bench lines for every single-GPU workload, ncu launch list and
# one `--set full` capture of the spread kernel.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/${tag}_smi.txt 2>&1
nproc >> $out/${tag}_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -3 $out/${tag}_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_c3_t1.json 2> $out/${tag}_bench_c3_t1.err
tail -c 1500 $out/${tag}_bench_c3_t1.json
for w in c3_t2 c2_t2 c2_t1 c4_t1 c1_t1; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err
done
for w in c3_t1 c3_t2; do
  timeout 400 python bench.py --workload $w --dist cluster --steps 10 --warmup 3 --no-cpu > $out/${tag}_bench_${w}_cluster.json 2> $out/${tag}_bench_${w}_cluster.err
done
timeout 300 python tools/bench_type3.py > $out/${tag}_bench_c5_t3.txt 2>&1
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
# launch list of the bench command (per-launch times are cold-cache; the SHARE is what counts)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $out/${tag}_launches_c3_t1.csv python bench.py --steps 2 --warmup 3 --no-cpu > $out/${tag}_ncu_bench.log 2>&1
# full capture of the dominant kernels (one launch each)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep3 -c 2 -f \
  -o $out/${tag}_sweep_spread python tools/prof_run.py --workload c3_t1 --reps 2 > $out/${tag}_ncu_full_spread.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep3 -c 1 -f \
  -o $out/${tag}_sweep_interp python tools/prof_run.py --workload c3_t2 --reps 1 > $out/${tag}_ncu_full_interp.log 2>&1
# the 2D sweep kernels (config C2)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep2 -c 1 -f \
  -o $out/${tag}_sweep2_interp python tools/prof_run.py --workload c2_t2 --reps 1 > $out/${tag}_ncu_full_interp2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep2 -c 1 -f \
  -o $out/${tag}_sweep2_spread python tools/prof_run.py --workload c2_t1 --reps 1 > $out/${tag}_ncu_full_spread2.log 2>&1
timeout 400 python bench.py --workload c2_t1 --dist cluster --steps 5 --warmup 3 --no-cpu > $out/${tag}_bench_c2_t1_cluster.json 2> $out/${tag}_bench_c2_t1_cluster.err
ls -la $out
