#!/bin/bash
# One GPU-box visit: parity tests (one process per file so a faulting kernel cannot poison the rest),
# smoke, a short bench, and the ncu launch list of the bench command.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
tag=${1:-run}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
for f in tests/test_gpu_csr.py tests/test_gpu_kernels.py tests/test_gpu_tiled.py tests/test_gpu_model.py tests/test_gpu_heads.py tests/test_gpu_tc.py tests/test_gpu_arena.py tests/test_gpu_fullsize.py tests/test_gpu_attribution.py tests/test_lite.py tests/test_viz.py tests/test_finetune_loop.py; do
  timeout 600 python -m pytest $f -q --no-header -p no:cacheprovider -m gpu 2>&1 | tail -120 > gpurun_out/${tag}_$(basename $f .py).log
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --precision fp32 --no-cpu-baseline > gpurun_out/${tag}_bench_fp32.log 2>&1
if [ "${2:-}" = "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-roofline \
    > gpurun_out/${tag}_ncu_bench.log 2>&1
fi
for f in gpurun_out/${tag}_test_*.log gpurun_out/${tag}_smoke.log; do echo "== $f"; tail -n 4 $f; done
tail -n 2 gpurun_out/${tag}_bench.log gpurun_out/${tag}_bench_fp32.log
if [ "${3:-}" = "full" ]; then
  # one `ncu --set full` capture of the message-passing kernels (bond-graph launches of layer >= 1)
  for k in k_gat_fwd_tiled k_gat_bwd_dst_tiled k_gat_bwd_src_tiled; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 0 -c 3 -f \
      -o gpurun_out/${tag}_full_$k python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-roofline \
      > gpurun_out/${tag}_full_$k.log 2>&1
  done
fi
