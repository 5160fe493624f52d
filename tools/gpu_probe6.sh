#!/bin/bash
tag=${1:-p}; out=gpurun_out
timeout 150 python bench.py --workload c3_t1 --steps 10 --warmup 3 --no-cpu --no-extras > $out/${tag}_c3.json 2>$out/${tag}_c3.err
python - $out/${tag}_c3.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'], 'stages', d['stages_ms'], 'acc', d['accuracy']['relerr'])
PY
timeout 250 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference.py -q -k "sweep3d or transform_matches or reproduces_reference or c3_grid" 2>&1 | tail -2
