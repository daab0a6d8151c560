#!/bin/bash
# Dev helper run under gpurun.  Output -> gpurun_out/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_all.log
tail -14 gpurun_out/pytest_gpu_all.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; cut -c1-1500 gpurun_out/bench_final.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1
grep -c "scan_lb_kernel" gpurun_out/launches_bench.csv
