#!/bin/bash
# Lean multi-GPU visit (large N costs N x box time): gpurun --gpus N --timeout 420 -- 'bash tools/gpu_multi_lean.sh tag N'
tag=${1:-m}
n=${2:-8}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29631"
run() {
  name=$1; to=$2; shift 2
  timeout $to $TR bench.py --gpus $n "$@" > $out/${tag}_bench_${name}_g$n.json 2> $out/${tag}_bench_${name}_g$n.err
  echo "== $name exit $?"; tail -c 1800 $out/${tag}_bench_${name}_g$n.json; grep -E "Error|error|failed" $out/${tag}_bench_${name}_g$n.err | tail -3
}
run c3_t1 110 --steps 10 --warmup 3 --workload c3_t1
run c3_t1_cluster 70 --steps 10 --warmup 3 --workload c3_t1 --dist cluster --no-extras
run c3_t2 70 --steps 10 --warmup 3 --workload c3_t2 --no-extras
run c4_t1 100 --steps 5 --warmup 3 --workload c4_t1 --no-cpu
