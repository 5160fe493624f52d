#!/bin/bash
# parity tests + a few bench lines:  gpurun -- 'bash tools/gpu_quick.sh tag "c3_t2 c3_t1"'
out=gpurun_out; mkdir -p $out
tag=${1:-q}; wl=${2:-"c3_t1 c3_t2"}
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
for w in $wl; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > $out/${tag}_${w}.json 2> $out/${tag}_${w}.err
  python - <<PY
import json
try:
    d=json.loads(open("$out/${tag}_${w}.json").read().strip().splitlines()[-1])
    print("$w", "ms/step %.3f"%d["ms_per_step"], {k: round(v,3) for k,v in d["stages_ms"].items()}, "setpts %.2f"%d["setpts_ms"], "e2e %.4g pts/s  %.2f ms"%(d["e2e"]["value"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print("$w FAILED", e)
PY
done
