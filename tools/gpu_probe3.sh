#!/bin/bash
tag=${1:-p}; out=gpurun_out
timeout 120 python tools/bench_type3.py > $out/${tag}_t3.txt 2>&1; tail -1 $out/${tag}_t3.txt | cut -c1-300
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_t3_launches.csv python tools/bench_type3.py 1e7 > $out/${tag}_t3_ncu.log 2>&1
timeout 120 python -m pytest tests/test_gpu_options.py -q -k auto_upsampfac 2>&1 | tail -2
