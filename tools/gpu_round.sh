#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "osa or golden or family or integer or single" > gpurun_out/pytest_osa.log 2>&1; tail -2 gpurun_out/pytest_osa.log
RF_CFG_SCALE=0.3 timeout 600 python tools/bench_configs.py mw > gpurun_out/cfg_mw.jsonl 2> gpurun_out/cfg_mw.err; tail -2 gpurun_out/cfg_mw.err; cut -c1-200 gpurun_out/cfg_mw.jsonl | tail -1
