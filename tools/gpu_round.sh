#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2w.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_r2w.log
python tools/bench_configs.py c5 2>/dev/null | cut -c1-300
