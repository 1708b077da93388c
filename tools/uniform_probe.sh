#!/bin/bash
# Compile tools/ubench_uniform.cu for sm_100a (no GPU needed) and count uniform-register vs vector-register FFMA2 operands per variant.
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cubin -o /tmp/ubench_uniform.cubin tools/ubench_uniform.cu
names=("mask from threadIdx.x >> 5, constant-bound loop" "mask from __shfl_sync(warp, 0)" "mask from %warpid" "tile loop starting at a warp-derived index"
       "uniform tile loop, guard if (t < n_tiles)" "uniform tile loop, guard through vote.all" "+ mbarrier spin with a per-thread exit"
       "+ mbarrier spin exiting on vote.all" "tcgen05.st under the warp-derived branch" "tcgen05.st outside any non-uniform branch")
echo "| variant | FFMA2 with a uniform-register operand | FFMA2 total | twiddle loads |"
echo "|---|---|---|---|"
for v in 0 1 2 3 4 5 6 7 8 9; do
  s=$(cuobjdump -sass -fun "_Z1kILi${v}EEvPfi" /tmp/ubench_uniform.cubin)
  ur=$(echo "$s" | grep -c 'FFMA2.*UR' || true); tot=$(echo "$s" | grep -c 'FFMA2' || true)
  ldcu=$(echo "$s" | grep -c 'LDCU.*c\[0x3\]' || true); ldc=$(echo "$s" | grep -c 'LDC\.64.*c\[0x3\]' || true)
  echo "| $v: ${names[$v]} | $ur | $tot | $ldcu LDCU / $ldc LDC |"
done
