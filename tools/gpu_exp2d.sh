#!/bin/bash
# 2D sweep kernels: parity tests, then A/B timing against the generic kernels.
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/x2d_pytest.log 2>&1; echo "pytest exit $?" >> $out/x2d_pytest.log
tail -5 $out/x2d_pytest.log
for w in c2_t2 c2_t1 c4_t1; do
  for sw in 1 0; do
    B200_NUFFT_SWEEP=$sw timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > $out/x2d_${w}_sw$sw.json 2> $out/x2d_${w}_sw$sw.err
    python - <<PY
import json
try:
    d=json.loads(open("$out/x2d_${w}_sw$sw.json").read().strip().splitlines()[-1])
    print("$w sweep=$sw", "ms/step %.3f"%d["ms_per_step"], d["stages_ms"], "setpts %.2f"%d["setpts_ms"], "frac %.3f"%d["roofline"]["frac"], "e2e %.3g"%d["e2e"]["value"])
except Exception as e:
    print("$w sweep=$sw FAILED", e); print(open("$out/x2d_${w}_sw$sw.err").read()[-1500:])
PY
  done
done
for it in 1024 2048 8192 16384; do
  for w in c2_t2 c2_t1; do
    B200_SWEEP2_ITEM=$it timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > $out/x2d_${w}_it$it.json 2>/dev/null
    python - <<PY
import json
try:
    d=json.loads(open("$out/x2d_${w}_it$it.json").read().strip().splitlines()[-1])
    print("$w item=$it", "ms/step %.3f"%d["ms_per_step"], d["stages_ms"], "setpts %.2f"%d["setpts_ms"])
except Exception as e:
    print("$w item=$it FAILED", e)
PY
  done
done
