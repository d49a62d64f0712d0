#!/bin/bash
# Dev tool: build a variant of libdqomap_b200.so with extra nvcc defines into dqo-map_b200/csrc/variants/<name>.so
#   tests/dev_variant.sh occ5 -DRB_OCC=5        then on the GPU:  DQO_B200_LIB=dqo-map_b200/csrc/variants/occ5.so python tests/dev_stage_times.py
set -e
name=$1; shift
cd "$(dirname "$0")/../dqo-map_b200/csrc"
mkdir -p variants/obj_$name
for f in *.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o variants/obj_$name/${f%.cu}.o &
done
wait
nvcc -shared -o variants/$name.so variants/obj_$name/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
rm -rf variants/obj_$name
echo variants/$name.so
