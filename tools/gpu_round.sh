#!/bin/bash
# Dev helper run under gpurun.  Output -> gpurun_out/
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_configs.py fam > gpurun_out/cfg_fam2.jsonl 2> gpurun_out/cfg_fam2.err
tail -3 gpurun_out/cfg_fam2.err; cut -c1-250 gpurun_out/cfg_fam2.jsonl
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/san_mem.log 2>&1; tail -4 gpurun_out/san_mem.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/san_race.log 2>&1; tail -4 gpurun_out/san_race.log
