#!/bin/bash
# Dev helper run under gpurun (edit per experiment).  Output -> gpurun_out/
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2c.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_r2c.log
python tools/bench_configs.py mw > gpurun_out/r2_mw_new2.jsonl 2> gpurun_out/r2_mw_new2.err; echo "mw new rc=$?"; cut -c1-330 gpurun_out/r2_mw_new2.jsonl
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_r2c.json'))
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'csr',d['e2e']['csr_u32']['value'],'ok',d['run'])
P
tail -3 gpurun_out/bench_r2c.err
RF_CFG_SCALE=0.3 ncu --set full --clock-control none --import-source on -k regex:scan_lbn -s 3 -c 1 -o gpurun_out/r2_lbn2 python tools/bench_configs.py mw > gpurun_out/ncu_lbn2.log 2>&1; echo "ncu lbn rc=$?"
