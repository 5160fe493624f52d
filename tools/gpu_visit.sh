#!/bin/bash
# One bounded GPU visit: selected tests, then the type-3 automatic-upsampfac probe.
# gpurun --timeout 400 -- 'bash tools/gpu_visit.sh tag "pytest -k expr"'
tag=${1:-v}
out=gpurun_out
mkdir -p $out
timeout 200 python -m pytest tests -m gpu -q -s -x -k "$2" > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log
grep -E "auto|passed|failed|Error|error|exit" $out/${tag}_pytest.log | tail -30
timeout 120 python tools/t3_auto_probe.py > $out/${tag}_t3_auto.jsonl 2> $out/${tag}_t3_auto.err
cat $out/${tag}_t3_auto.jsonl; tail -3 $out/${tag}_t3_auto.err
