#!/bin/bash
mkdir -p gpurun_out
RF_CFG_SCALE=0.3 timeout 300 ncu --set full --import-source on --clock-control none -k regex:scan_jaro32 -s 3 -c 1 -f -o gpurun_out/j32 python tools/bench_configs.py c4 > gpurun_out/ncu_j32.log 2>&1
python tools/ncu_summary.py gpurun_out/j32.ncu-rep > gpurun_out/j32_summary.txt 2>&1; head -24 gpurun_out/j32_summary.txt
RF_CFG_SCALE=0.1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:cdist_scan -s 1 -c 1 -f -o gpurun_out/cdist python tools/bench_configs.py c5 > gpurun_out/ncu_cdist.log 2>&1
python tools/ncu_summary.py gpurun_out/cdist.ncu-rep > gpurun_out/cdist_summary.txt 2>&1; head -24 gpurun_out/cdist_summary.txt
