#!/bin/bash
# Multi-GPU visit: gpurun --gpus N --timeout 1200 -- 'bash tools/gpu_multi.sh tag N [big]'
tag=${1:-m}
n=${2:-2}
big=${3:-}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29631"
timeout 600 $TR tools/sharded_check.py $big > $out/${tag}_sharded_check_g$n.log 2>&1
echo "sharded_check exit $?" >> $out/${tag}_sharded_check_g$n.log
grep -E "sharded|Error|error" $out/${tag}_sharded_check_g$n.log | tail -30
for wl in c3_t1 c3_t2; do
  timeout 600 $TR bench.py --gpus $n --steps 10 --warmup 3 --workload $wl > $out/${tag}_bench_${wl}_g$n.json 2> $out/${tag}_bench_${wl}_g$n.err
  tail -c 3000 $out/${tag}_bench_${wl}_g$n.json; tail -5 $out/${tag}_bench_${wl}_g$n.err
done
timeout 600 $TR bench.py --gpus $n --steps 10 --warmup 3 --workload c3_t1 --dist cluster --no-extras > $out/${tag}_bench_c3_t1_cluster_g$n.json 2> $out/${tag}_bench_c3_t1_cluster_g$n.err
tail -c 2500 $out/${tag}_bench_c3_t1_cluster_g$n.json; tail -5 $out/${tag}_bench_c3_t1_cluster_g$n.err
timeout 900 $TR bench.py --gpus $n --steps 5 --warmup 3 --workload c4_t1 --no-cpu > $out/${tag}_bench_c4_t1_g$n.json 2> $out/${tag}_bench_c4_t1_g$n.err
tail -c 1500 $out/${tag}_bench_c4_t1_g$n.json; tail -5 $out/${tag}_bench_c4_t1_g$n.err
