#!/bin/bash
# Dev helper run under gpurun.  Output -> gpurun_out/
mkdir -p gpurun_out
RF_LIB_PATH=$PWD/rapidfuzz-rs_b200/lib/librfgpu_hparith.so timeout 120 python bench.py --steps 100 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_hparith.json 2> gpurun_out/bench_hparith.err
python -c "import json;d=json.load(open('gpurun_out/bench_hparith.json'));print('HP-arith variant: ms_per_step', d['ms_per_step'], 'matches oracle sample', d['config']['results_match_oracle_sample'])"
