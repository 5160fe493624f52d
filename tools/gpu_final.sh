#!/bin/bash
# what the driver runs at round end, on one GPU: smoke, the GPU test suite, the default bench
tag=${1:-fin}; out=gpurun_out; mkdir -p $out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke exit $?" | tee -a $out/${tag}_smoke.log; tail -3 $out/${tag}_smoke.log
timeout 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log; tail -4 $out/${tag}_pytest.log
timeout 300 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit $?"; tail -c 600 $out/${tag}_bench.json
