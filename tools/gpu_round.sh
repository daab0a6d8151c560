#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "single_word or long_cand" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_configs.py c4 > gpurun_out/cfg4.jsonl 2> gpurun_out/cfg4.err
cut -c1-250 gpurun_out/cfg4.jsonl; tail -3 gpurun_out/cfg4.err
