#!/bin/bash
tag=${1:-p}; out=gpurun_out
export B200_NUFFT_PRECOMP=1
timeout 200 python bench.py --workload c3_t1 --steps 10 --warmup 3 --no-cpu > $out/${tag}_c3.json 2>$out/${tag}_c3.err
python - $out/${tag}_c3.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('PRECOMP t1 ms/step', d['ms_per_step'], 'stages', d['stages_ms'], 'acc', d['accuracy']['relerr'], 'setpts', d['setpts_ms'], 't2', d.get('type2',{}).get('ms_per_step'), d.get('type2',{}).get('stages_ms'))
PY
tail -3 $out/${tag}_c3.err
timeout 250 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference.py -q -k "sweep3d or transform_matches or reproduces_reference or c3_grid or sort_ragged" 2>&1 | tail -2
