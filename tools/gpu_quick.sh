#!/bin/bash
# Quick GPU visit: selected tests + one bench line.  gpurun --timeout 900 -- 'bash tools/gpu_quick.sh tag [pytest -k expr]'
tag=${1:-q}
kexpr=${2:-}
out=gpurun_out
mkdir -p $out
if [ -n "$kexpr" ]; then
  timeout 1200 python -m pytest tests -m gpu -q -s -k "$kexpr" > $out/${tag}_pytest.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1
fi
echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -25 $out/${tag}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_c3_t1.json 2> $out/${tag}_bench_c3_t1.err
tail -c 2500 $out/${tag}_bench_c3_t1.json; tail -5 $out/${tag}_bench_c3_t1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
tail -c 1500 $out/${tag}_bench_reference.json; tail -5 $out/${tag}_bench_reference.err
