#!/bin/bash
# Dev helper run under gpurun.  Output -> gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cdist" --durations=3 > gpurun_out/pytest_cdist.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_cdist.log
tail -8 gpurun_out/pytest_cdist.log
: > gpurun_out/cfg5_v2.jsonl
for o in "cdist_skip=1" "cdist_skip=0" "cdist_skip=1,cdist_slices=2" "cdist_skip=1,cdist_slices=8"; do
  echo "# $o" >> gpurun_out/cfg5_v2.jsonl
  RF_OPTS=$o timeout 300 python tools/bench_configs.py c5 >> gpurun_out/cfg5_v2.jsonl 2>> gpurun_out/cfg5_v2.err
done
cat gpurun_out/cfg5_v2.jsonl
