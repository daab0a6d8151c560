#!/bin/bash
mkdir -p gpurun_out
for v in "" _RF_ROW_LDG _RF_ROW_LDCG _RF_ROW_LU _RF_PF_L1 _RF_PF_EL; do
  RF_LIB_PATH=$PWD/rapidfuzz-rs_b200/lib/librfgpu$v.so timeout 300 python bench.py --steps 200 --no-cpu-baseline --e2e-steps 1 > gpurun_out/b.json 2> gpurun_out/b.err
  echo "variant [$v]: $(python -c "import json;d=json.load(open('gpurun_out/b.json'));print(d['ms_per_step'], d['config']['results_match_oracle_sample'])")"
done
