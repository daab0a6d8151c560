#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q --durations=5 > gpurun_out/pytest_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_full.log
tail -15 gpurun_out/pytest_full.log
free -g | head -2
