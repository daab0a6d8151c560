#!/bin/bash
# Dev helper run under gpurun (edit per experiment).  Output -> gpurun_out/
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py -x -q > gpurun_out/pytest_r2b.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_r2b.log
python tools/bench_configs.py mw > gpurun_out/r2_mw_new.jsonl 2> gpurun_out/r2_mw_new.err; echo "mw new rc=$?"; cut -c1-400 gpurun_out/r2_mw_new.jsonl
RF_OPTS=multi_word_path=1 python tools/bench_configs.py mw > gpurun_out/r2_mw_old.jsonl 2> gpurun_out/r2_mw_old.err; echo "mw old rc=$?"; cut -c1-300 gpurun_out/r2_mw_old.jsonl
RF_CFG_SCALE=0.3 ncu --set full --clock-control none --import-source on -k regex:scan_lbn -s 3 -c 1 -o gpurun_out/r2_lbn python tools/bench_configs.py mw > gpurun_out/ncu_lbn.log 2>&1; echo "ncu lbn rc=$?"
RF_CFG_SCALE=0.3 RF_OPTS=multi_word_path=1 ncu --set full --clock-control none --import-source on -k regex:scan_mw -s 3 -c 1 -o gpurun_out/r2_mw_old python tools/bench_configs.py mw > gpurun_out/ncu_mw.log 2>&1; echo "ncu mw rc=$?"
