# Repeats one test under the stream / PDL switches to localise a scheduling-dependent failure.
set -u
t=${1:-tests/test_gpu_heads.py::test_pretrain_model_training_step_matches_oracle}
for mode in "" "FNB_PDL=0" "FNB_STREAMS=1"; do
  for i in 1 2 3 4 5 6; do
    env $mode timeout 300 python -m pytest "$t" -q --no-header -p no:cacheprovider -m gpu 2>&1 | grep -E "passed|failed|AssertionError: \{" | tr '\n' ' '
    echo " [$mode #$i]"
  done
done
