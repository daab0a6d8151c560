#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 1200 gpurun_out/bench_final.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cut -c1-200 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_lb_kernel -s 3 -c 1 -f -o gpurun_out/prof_lb_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_lb.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:band_ -s 3 -c 3 -f -o gpurun_out/prof_band_final python tools/bench_configs.py c3 > gpurun_out/ncu_band.log 2>&1
timeout 600 python tools/bench_configs.py c3,c4 > gpurun_out/cfg34.jsonl 2>/dev/null
cat gpurun_out/cfg34.jsonl | cut -c1-260
ls -la gpurun_out/*.ncu-rep
