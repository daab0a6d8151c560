#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "cdist" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_configs.py c5 > gpurun_out/cfg5.jsonl 2> gpurun_out/cfg5.err
cat gpurun_out/cfg5.jsonl; tail -3 gpurun_out/cfg5.err
