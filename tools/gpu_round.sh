#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hamming or golden or simple or prefix or compaction or cpp" > gpurun_out/pytest_ham.log 2>&1; tail -3 gpurun_out/pytest_ham.log
timeout 600 python tools/bench_configs.py simple > gpurun_out/cfg_simple2.jsonl 2> gpurun_out/cfg_simple2.err; tail -2 gpurun_out/cfg_simple2.err; cut -c1-300 gpurun_out/cfg_simple2.jsonl
