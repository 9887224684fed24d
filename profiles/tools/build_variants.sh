#!/bin/bash
# usage: build_variants.sh name[:nvcc flags] ...   -> build/variants/<name>.so from the current sources (in parallel)
mkdir -p build/variants
for v in "$@"; do
  n=${v%%:*}; f=""; [[ "$v" == *:* ]] && f=${v#*:}
  ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $f -Xptxas -v \
      -o build/variants/$n.so tactics2d_b200/csrc/t2d_kernels.cu > build/variants/$n.log 2>&1 \
    && echo "$n: $(grep -A2 'step_kernelILi4ELb1' build/variants/$n.log | grep -E 'registers' | sed 's/ptxas info    : //')" \
    || { echo "$n: BUILD FAILED"; grep error build/variants/$n.log | head -5; } ) &
done
wait
