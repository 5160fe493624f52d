#!/bin/bash
tag=${1:-p}; out=gpurun_out
for d in cluster uniform; do
timeout 150 python bench.py --workload c3_t1 --dist $d --steps 5 --warmup 3 --no-cpu --no-extras > $out/${tag}_c3_$d.json 2>$out/${tag}_c3_$d.err
python - $out/${tag}_c3_$d.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d['config']['workload'][-40:], 'setpts', d['setpts_ms'], 'ms/step', d['ms_per_step'], 'acc', d['accuracy']['relerr'])
PY
done
timeout 150 python bench.py --workload c2_t2 --dist cluster --steps 5 --warmup 3 --no-cpu --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 cluster setpts', d['setpts_ms'])"
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -q -k "sort_permutation or sort_ragged or full_size or c3_256 or many_points_in_one_bin" 2>&1 | tail -2
