#!/bin/bash
mkdir -p gpurun_out
python tools/dbg_jaro.py
timeout 300 python tools/bench_configs.py c4 > gpurun_out/cfg4_b.jsonl 2> gpurun_out/cfg4_b.err; cut -c1-260 gpurun_out/cfg4_b.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "jaro or golden or family or metric or config4 or u32" > gpurun_out/pytest_jaro.log 2>&1; tail -3 gpurun_out/pytest_jaro.log
