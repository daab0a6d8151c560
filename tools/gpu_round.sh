#!/bin/bash
# Dev helper run under gpurun: smoke + the full GPU test suite.  Output -> gpurun_out/
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_configs.py dp > gpurun_out/cfg_dp.jsonl 2> gpurun_out/cfg_dp.err
cat gpurun_out/cfg_dp.jsonl; tail -3 gpurun_out/cfg_dp.err
