#!/bin/bash
# compute-sanitizer passes over small parity cases (race / memory checks).
#   gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh [tag]'
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
SEL="test_transform_matches_oracle or test_sweep2d_ragged_and_batched or test_spreadinterp_only or test_no_sort_option_device_api"
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 --target-processes all \
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_options.py -x -q -k "$SEL" > $out/${tag}_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> $out/${tag}_sanitizer_memcheck.log
timeout 700 compute-sanitizer --tool racecheck --error-exitcode 9 --target-processes all \
  python -m pytest tests/test_gpu_parity.py -x -q -k "test_transform_matches_oracle" > $out/${tag}_sanitizer_racecheck.log 2>&1
echo "racecheck exit $?" >> $out/${tag}_sanitizer_racecheck.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sharded_check.py > $out/${tag}_sanitizer_memcheck_sharded.log 2>&1
echo "memcheck sharded exit $?" >> $out/${tag}_sanitizer_memcheck_sharded.log
for f in memcheck racecheck memcheck_sharded; do
  echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|exit|passed|failed" $out/${tag}_sanitizer_$f.log | tail -6
done
