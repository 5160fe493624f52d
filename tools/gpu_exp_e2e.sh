#!/bin/bash
out=gpurun_out; mkdir -p $out
tag=${1:-e2e}
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -4 $out/${tag}_pytest.log
show() {
python - <<PY
import json
try:
    d=json.loads(open("$1").read().strip().splitlines()[-1])
    print("$2", "ms/step %.3f"%d["ms_per_step"], {k: round(v,3) for k,v in d["stages_ms"].items()}, "setpts %.2f"%d["setpts_ms"], "e2e %.4g pts/s  %.2f ms"%(d["e2e"]["value"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print("$2 FAILED", e)
PY
}
for w in c3_t1 c3_t2 c2_t2 c4_t1; do
  for g in 1 2 4 8; do
    if [ $w = c4_t1 ] && [ $g != 1 ]; then continue; fi
    B200_NUFFT_HOST_GROUPS=$g timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > $out/${tag}_${w}_g$g.json 2> $out/${tag}_${w}_g$g.err
    show $out/${tag}_${w}_g$g.json "$w groups=$g"
  done
done
