#!/bin/bash
# Dev helper run under gpurun.  Output -> gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=4 > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_all.log
tail -9 gpurun_out/pytest_gpu_all.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; cut -c1-300 gpurun_out/bench_final.json
