#!/bin/bash
# build a kernel-variant library for A/B probes: tools/build_variant.sh NAME "-DMACRO=.. -DMACRO=.."
name=$1; defs=$2
B=finufft_b200/build; V=finufft_b200/variants; C=finufft_b200/csrc
mkdir -p $V/obj_$name
FL="-gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
for f in partition engine; do
  /usr/local/cuda/bin/nvcc $FL $defs -c $C/$f.cu -o $V/obj_$name/$f.o &
done
wait
objs=$(ls $B/*.o | grep -v -E "/(partition|engine)\.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $V/lib_$name.so $objs $V/obj_$name/partition.o $V/obj_$name/engine.o -lcufft -ldl
ls -la $V/lib_$name.so
