#!/bin/bash
mkdir -p gpurun_out
for v in "" _lb512; do
  RF_LIB_PATH=$PWD/rapidfuzz-rs_b200/lib/librfgpu$v.so timeout 300 python bench.py --steps 200 --no-cpu-baseline --e2e-steps 1 > gpurun_out/b.json 2> gpurun_out/b.err
  echo "variant [$v]: $(python -c "import json;d=json.load(open('gpurun_out/b.json'));print(d['ms_per_step'], d['config']['results_match_oracle_sample'])")"
done
RF_LIB_PATH=$PWD/rapidfuzz-rs_b200/lib/librfgpu_lb512.so timeout 300 ncu --metrics launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,smsp__warps_active.avg.per_cycle_active,smsp__inst_executed.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:scan_lb -s 3 -c 1 --csv --log-file gpurun_out/lb_occ.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > /dev/null 2>&1
cut -d, -f13- gpurun_out/lb_occ.csv | tail -6
