#!/bin/bash
# launch list of one setpts + execute with the current library: gpurun --timeout 400 -- 'bash tools/gpu_probe2.sh tag'
tag=${1:-p}
out=gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
  --log-file $out/${tag}_launches_setpts.csv python tools/prof_run.py --workload c3_t1 --reps 1 > $out/${tag}_ncu_setpts.log 2>&1
tail -2 $out/${tag}_ncu_setpts.log | cut -c1-200
timeout 120 python -m pytest tests/test_gpu_parity.py -q -x -k "sort_permutation or sort_ragged or full_size" 2>&1 | tail -3
timeout 100 python tools/sharded_check.py --big 2>&1 | grep -E "sharded_check|FAIL" | tail -3
