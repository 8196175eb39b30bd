#!/bin/bash
# `ncu --set full` captures (with source correlation) of the bond-graph launches of the three message-passing kernels
# inside one bench step; reports land in gpurun_out/<tag>_full_<kernel>.ncu-rep.
set -u
tag=${1:-run}
mkdir -p gpurun_out
# launch order inside one fused step: forward bond first; backward frag, fragment-connection, atom, bond
for spec in "k_gat_fwd_tiled 0" "k_gat_bwd_dst_tiled 3" "k_gat_bwd_src_tiled 3"; do
  set -- $spec
  FNB_STREAMS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f \
    -o gpurun_out/${tag}_full_$1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-roofline \
    > gpurun_out/${tag}_full_$1.log 2>&1
done
ls -la gpurun_out/${tag}_full_*
