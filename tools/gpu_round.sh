#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2l.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_r2l.log
python tools/bench_configs.py mw 2>gpurun_out/r2_mw_new4.err | head -3 > gpurun_out/r2_mw_new4.jsonl; cut -c1-330 gpurun_out/r2_mw_new4.jsonl
