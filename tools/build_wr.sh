#!/bin/bash
# usage: build_wr.sh name "flags" -> tools/wr_<name>
cd /root/repo/tools && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false --expt-relaxed-constexpr -DEXPNAME="\"$1\"" $2 -o wr_$1 walker_real.cu 2>&1 | grep -i -E "error" 
