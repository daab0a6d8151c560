#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "damerau or golden or cpp or hamming" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_configs.py dp > gpurun_out/cfg_dp.jsonl 2> gpurun_out/cfg_dp.err
cat gpurun_out/cfg_dp.jsonl; tail -3 gpurun_out/cfg_dp.err
