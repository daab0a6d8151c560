#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:band_run -s 2 -c 2 -f -o gpurun_out/prof_band python tools/bench_configs.py c3 > gpurun_out/ncu_band.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_lb -s 3 -c 1 -f -o gpurun_out/prof_lb python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_lb.log 2>&1
ls -la gpurun_out/*.ncu-rep
