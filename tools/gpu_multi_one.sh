#!/bin/bash
# Headline multi-GPU line only: gpurun --gpus N --timeout 200 -- 'bash tools/gpu_multi_one.sh tag N'
tag=${1:-m}; n=${2:-8}; out=gpurun_out; mkdir -p $out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29631 \
  bench.py --gpus $n --steps 10 --warmup 3 --workload c3_t1 > $out/${tag}_bench_c3_t1_g$n.json 2> $out/${tag}_bench_c3_t1_g$n.err
echo "exit $?"; tail -c 2500 $out/${tag}_bench_c3_t1_g$n.json; grep -E "Error|error|failed" $out/${tag}_bench_c3_t1_g$n.err | tail -3
