#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_jaro32 -s 2 -c 1 -f -o gpurun_out/prof_j32 python tools/bench_configs.py c4 > gpurun_out/ncu_j32.log 2>&1
