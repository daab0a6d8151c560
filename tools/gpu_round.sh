#!/bin/bash
# Dev helper run under gpurun.  Output -> gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cdist" --durations=3 > gpurun_out/pytest_cdist.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_cdist.log
tail -8 gpurun_out/pytest_cdist.log
timeout 600 python tools/bench_cdist_sharded.py > gpurun_out/cdist_sharded_n1.json 2> gpurun_out/cdist_sharded_n1.err
tail -3 gpurun_out/cdist_sharded_n1.err; cat gpurun_out/cdist_sharded_n1.json
