set -e; python -c "from fragnet_b200 import _abi; _abi.load()"; set +e
set -u
tag=$1
mkdir -p gpurun_out
for f in tests/test_gpu_tiled.py tests/test_gpu_model.py tests/test_gpu_heads.py; do
  timeout 600 python -m pytest $f -q --no-header -p no:cacheprovider -m gpu 2>&1 | tail -60 > gpurun_out/${tag}_$(basename $f .py).log
  tail -n 3 gpurun_out/${tag}_$(basename $f .py).log
done
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench.log 2>gpurun_out/${tag}_bench.err
tail -n 1 gpurun_out/${tag}_bench.log | cut -c1-400
FNB_SEQ=1 timeout 300 python scripts/device_profile.py > gpurun_out/${tag}_device_profile.log 2>&1
grep "device busy" gpurun_out/${tag}_device_profile.log
