#!/bin/bash
tag=${1:-p}; n=${2:-2}; out=gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29641 tools/bench_type3.py > $out/${tag}_t3_g$n.txt 2> $out/${tag}_t3_g$n.err
tail -1 $out/${tag}_t3_g$n.txt | cut -c1-400; grep -E "Error|error" $out/${tag}_t3_g$n.err | tail -3
