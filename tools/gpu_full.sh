#!/bin/bash
# Full GPU test suite + ncu launch list of one setpts+execute.  gpurun --timeout 2400 -- 'bash tools/gpu_full.sh tag'
tag=${1:-f}
out=gpurun_out
mkdir -p $out
timeout 2000 python -m pytest tests -m gpu -q -s > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log
grep -E "passed|failed|FAILED|Error|gpu-vs" $out/${tag}_pytest.log | tail -40
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
  --log-file $out/${tag}_launches_setpts.csv python tools/prof_run.py --workload c3_t1 --reps 1 > $out/${tag}_ncu_setpts.log 2>&1
tail -3 $out/${tag}_ncu_setpts.log
