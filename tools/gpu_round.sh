#!/bin/bash
# Dev helper run under gpurun.  Output -> gpurun_out/
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_configs.py c2w > gpurun_out/cfg_c2w.jsonl 2> gpurun_out/cfg_c2w.err
tail -3 gpurun_out/cfg_c2w.err; cat gpurun_out/cfg_c2w.jsonl
