#!/bin/bash
# Dev helper run under gpurun.  Output -> gpurun_out/
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python tools/bench_configs.py widen > gpurun_out/cfg_widen2.jsonl 2> gpurun_out/cfg_widen2.err
tail -3 gpurun_out/cfg_widen2.err; cut -c1-400 gpurun_out/cfg_widen2.jsonl
