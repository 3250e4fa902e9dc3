#!/bin/bash
# A/B builds of one kernel file: tests/tools/ab_build.sh NAME FILE.cu [-DFLAG=...]...
# compiles csrc/FILE.cu with the extra flags and links it with the other objects of the regular build into
# omok-ai_b200/_build/variants/NAME.so (travels to the GPU box; select it there with `cp` over libomok_b200.so).
set -e
cd "$(dirname "$0")/../../omok-ai_b200/csrc"
name=$1; file=$2; shift 2
make -s >/dev/null
mkdir -p ../_build/variants
base=$(basename "$file" .cu)
extra=""
[ "$base" = tree_kernels ] && extra="-fmad=false"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden $extra "$@" -c "$file" -o ../_build/variants/$name.$base.o
objs=""
for o in env_kernels tree_kernels net_kernels fc_f16 tower_f16 train_kernels omk_api; do
  if [ "$o" = "$base" ]; then objs="$objs ../_build/variants/$name.$base.o"; else objs="$objs ../_build/$o.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../_build/variants/$name.so $objs -lcudart -ldl
echo built ../_build/variants/$name.so
