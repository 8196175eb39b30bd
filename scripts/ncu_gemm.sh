#!/bin/bash
# `ncu --set full` captures of the tensor-core projection kernels (forward / dX: k_tc_proj, k_tc_proj3r; weight gradient:
# k_tc_dw, k_tc_dw3) on a bond-graph-sized operand (53 940 x 128), L2 flushed before the captured launch.
# Reports land in gpurun_out/<tag>_full_<kernel>.ncu-rep; summarise with scripts/summarize_ncu_full.py.
set -u
tag=${1:-run}
mkdir -p gpurun_out
for k in k_tc_proj3r k_tc_proj\( k_tc_dw3 k_tc_dw\(; do
  name=$(echo $k | tr -d '\\(')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 4 -c 1 -f \
    -o gpurun_out/${tag}_full_${name} python scripts/gemm_bench.py --rows 53940 --iters 2 \
    > gpurun_out/${tag}_full_${name}.log 2>&1
done
ls -la gpurun_out/${tag}_full_*
