#!/bin/bash
# Builds variants of libgridpp_b200.so with different -D flags into scratch/ for A/B timing on the GPU box:
#   profiles/variants.sh name1 "-DNBH_PF=1 -DNBH_MINB_SUM=4" name2 "..."
# then on the box: GPP_B200_LIB=$PWD/scratch/lib_name1.so python profiles/nbh_time.py
set -e
cd "$(dirname "$0")/.."
mkdir -p scratch
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fopenmp,-O2 -shared \
    $flags -I include -I gridpp_b200/csrc -o scratch/lib_$name.so gridpp_b200/csrc/{capi,points,oi,neighbourhood,neighbourhood_tma,quantile_tma,ensemble_forms,thresholds,ensi,stats,gridding,multi_gpu,ensi_multi}.cu -lgomp &
done
wait
ls -la scratch/*.so
