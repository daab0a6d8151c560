#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_configs.py post > gpurun_out/cfg_post.jsonl 2> gpurun_out/cfg_post.err
cat gpurun_out/cfg_post.jsonl; tail -3 gpurun_out/cfg_post.err
