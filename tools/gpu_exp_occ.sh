#!/bin/bash
out=gpurun_out; mkdir -p $out
tag=${1:-occ}
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
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
for w in c3_t1 c3_t2; do
  for o in 16 20 24; do
    B200_SWEEP3_OCC=$o timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > $out/${tag}_${w}_o$o.json 2> $out/${tag}_${w}_o$o.err
    show $out/${tag}_${w}_o$o.json "$w occ=$o"
  done
done
for o in 16 20 24; do
  B200_SWEEP3_OCC=$o timeout 300 python bench.py --workload c3_t1 --dist cluster --steps 5 --warmup 3 --no-cpu > $out/${tag}_c3_t1_cl_o$o.json 2>/dev/null
  show $out/${tag}_c3_t1_cl_o$o.json "c3_t1 cluster occ=$o"
done
for w in c4_t1 c2_t1 c2_t2; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > $out/${tag}_${w}.json 2> $out/${tag}_${w}.err
  show $out/${tag}_${w}.json "$w"
done
