#!/bin/bash
# One bounded GPU visit: selected tests, then the point-group layout probe.
# gpurun --timeout 400 -- 'bash tools/gpu_visit.sh tag "pytest -k expr" "type-1 cases" "type-2 cases"'
tag=${1:-v}
out=gpurun_out
mkdir -p $out
timeout 200 python -m pytest tests -m gpu -q -s -x -k "$2" > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log
grep -E "auto|passed|failed|Error|error|exit" $out/${tag}_pytest.log | tail -30
timeout 150 python tools/e2e_groups.py --workload c3_t1 --steps 8 --t1 "$3" --t2 "$4" > $out/${tag}_groups_c3.jsonl 2> $out/${tag}_groups_c3.err
cat $out/${tag}_groups_c3.jsonl; tail -3 $out/${tag}_groups_c3.err
