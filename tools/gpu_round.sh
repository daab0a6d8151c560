#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "damerau or golden or compaction or dp_metrics or weight or family or cpp" > gpurun_out/pytest_dl.log 2>&1; tail -3 gpurun_out/pytest_dl.log
timeout 600 python tools/bench_configs.py dp > gpurun_out/cfg_dp2.jsonl 2> gpurun_out/cfg_dp2.err; tail -2 gpurun_out/cfg_dp2.err; cut -c1-330 gpurun_out/cfg_dp2.jsonl
