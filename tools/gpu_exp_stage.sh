#!/bin/bash
# staged strength permutation: parity tests, then A/B timing (B200_NUFFT_STAGE=0/1)
out=gpurun_out; mkdir -p $out
tag=${1:-stg}
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -4 $out/${tag}_pytest.log
show() {
python - <<PY
import json
try:
    d=json.loads(open("$1").read().strip().splitlines()[-1])
    print("$2", "ms/step %.3f"%d["ms_per_step"], {k: round(v,3) for k,v in d["stages_ms"].items()}, "setpts %.2f"%d["setpts_ms"], "frac %.3f"%d["roofline"]["frac"], "e2e %.3g"%d["e2e"]["value"])
except Exception as e:
    print("$2 FAILED", e)
PY
}
for w in c2_t2 c2_t1 c3_t1 c3_t2 c4_t1; do
  for sg in 1 0; do
    B200_NUFFT_STAGE=$sg timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > $out/${tag}_${w}_sg$sg.json 2> $out/${tag}_${w}_sg$sg.err
    show $out/${tag}_${w}_sg$sg.json "$w stage=$sg"
  done
done
