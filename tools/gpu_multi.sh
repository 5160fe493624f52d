#!/bin/bash
# Multi-GPU visit (bounded): gpurun --gpus N --timeout 800 -- 'bash tools/gpu_multi.sh tag N [lean]'
tag=${1:-m}
n=${2:-2}
lean=${3:-}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29631"
run() {  # name, timeout, args...
  name=$1; to=$2; shift 2
  timeout $to $TR bench.py --gpus $n "$@" > $out/${tag}_bench_${name}_g$n.json 2> $out/${tag}_bench_${name}_g$n.err
  echo "== $name exit $?"; tail -c 2200 $out/${tag}_bench_${name}_g$n.json; grep -E "Error|error|failed" $out/${tag}_bench_${name}_g$n.err | tail -3
}
if [ "$lean" != lean ]; then
  timeout 150 $TR tools/sharded_check.py > $out/${tag}_sharded_check_g$n.log 2>&1
  echo "sharded_check exit $?" >> $out/${tag}_sharded_check_g$n.log
  grep -E "sharded_check|FAIL|exit" $out/${tag}_sharded_check_g$n.log | tail -5
fi
run c3_t1 200 --steps 10 --warmup 3 --workload c3_t1
run c3_t2 150 --steps 10 --warmup 3 --workload c3_t2 --no-extras
run c3_t1_cluster 150 --steps 10 --warmup 3 --workload c3_t1 --dist cluster --no-extras
[ "$lean" = noc4 ] || run c4_t1 200 --steps 5 --warmup 3 --workload c4_t1 --no-cpu
