#!/bin/bash
# Last GPU visit of a round on a short budget: the whole GPU suite on three workers (one test
# file per worker at a time), then one bench line.  gpurun --timeout 230 -- 'bash tools/gpu_last.sh tag'
tag=${1:-last}
out=gpurun_out
mkdir -p $out
timeout ${2:-150} python -m pytest tests -m gpu -q -x -n 3 --dist loadfile > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -8 $out/${tag}_pytest.log
timeout 60 python bench.py --steps 10 --warmup 3 --no-cpu > $out/${tag}_bench_c3_t1.json 2> $out/${tag}_bench_c3_t1.err
tail -c 1800 $out/${tag}_bench_c3_t1.json; tail -3 $out/${tag}_bench_c3_t1.err
