#!/bin/bash
# Dev helper run under gpurun.  Output -> gpurun_out/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "4gib" --durations=3 > gpurun_out/pytest_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_full.log
tail -25 gpurun_out/pytest_full.log
