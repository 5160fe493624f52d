#!/bin/bash
# short, bounded probes of kernel variants: gpurun --timeout 500 -- 'bash tools/gpu_probe.sh tag "A B C"'
tag=${1:-p}
out=gpurun_out
mkdir -p $out
for v in base $2; do
  if [ $v = base ]; then unset B200_NUFFT_LIBRARY; else export B200_NUFFT_LIBRARY=$PWD/finufft_b200/variants/lib_$v.so; fi
  timeout 60 python tools/prof_run.py --workload c3_t1 --reps 2 --setpts-ms > $out/${tag}_probe_$v.log 2>&1; echo "exit $?" >> $out/${tag}_probe_$v.log
  echo "== $v: $(tail -2 $out/${tag}_probe_$v.log | cut -c1-200 | tr '\n' ' ')"
done
