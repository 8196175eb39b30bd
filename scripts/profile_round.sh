#!/bin/bash
# Round profile of the fp32-mode bench step (GPU box, one GPU):
#   1. launch list: every kernel of `bench.py --steps 2 --warmup 1` with its device time (cold-cache, serialised)
#   2. `ncu --set full` of one launch each of the dominant kernels inside the step (FNB_STREAMS=1: single-stream order):
#      bond-graph attention forward / destination pass / source pass, the 3xTF32 projection and weight-gradient GEMMs,
#      and their TF32 counterparts (tensor-pipe utilisation, DRAM bytes, issue slots)
# Outputs in gpurun_out/<tag>_*; summarise here with scripts/summarize_launches.py, summarize_ncu_full.py,
# kernel_shares.py.
set -u
tag=${1:-r4}
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e --no-roofline --no-extras --no-tf32"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
  $B --steps 2 --warmup 1 > gpurun_out/${tag}_launches.log 2>&1
# kernel regex, launches to skip (bond graph = the largest launch of its kind inside one single-stream step)
for spec in "k_gat_fwd_tiled 0 1" "k_gat_bwd_dst_tiled 3 1" "k_gat_bwd_src_tiled 3 1" "k_tc_proj3r 4 3" "k_tc_dw3 2 3"; do
  set -- $spec
  FNB_STREAMS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -f \
    -o gpurun_out/${tag}_full_$1 $B --steps 1 --warmup 1 > gpurun_out/${tag}_full_$1.log 2>&1
done
for spec in "k_tc_proj 4 3" "k_tc_dw 2 3"; do
  set -- $spec
  FNB_STREAMS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^$1\$|$1\(" -s $2 -c $3 -f \
    -o gpurun_out/${tag}_full_$1_tf32 $B --precision tf32 --steps 1 --warmup 1 > gpurun_out/${tag}_full_$1_tf32.log 2>&1
done
ls -la gpurun_out/${tag}_full_* gpurun_out/${tag}_launches.csv
