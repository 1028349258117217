#!/bin/bash
# Build walker variants into kvmatch_b200/variants/<name>.so ; usage: tools/variants.sh name "EXTRA flags" ...
set -e
cd /root/repo/kvmatch_b200/csrc
mkdir -p ../variants
while [ $# -gt 1 ]; do
  name=$1; extra=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a $extra -O3 -std=c++17 -lineinfo -fmad=false --expt-relaxed-constexpr \
    -Xcompiler -fPIC,-O2,-ffp-contract=off,-fno-fast-math -Xptxas -v -shared -o ../variants/$name.so kvmatch_gpu.cu 2> ../variants/$name.log
  if cuobjdump -sass ../variants/$name.so | grep -q 'LDGSTS.*+UR0'; then echo "$name: bad LDGSTS"; fi
  grep -A2 "cnsm_relay_kernelILi[0-9]*ELi1ELi0" ../variants/$name.log | grep Used | sed "s/^/$name: /"
done
