#!/bin/bash
mkdir -p gpurun_out
for ch in 16 4 8 32 64 128; do
  RF_LB_CHUNK=$ch timeout 300 python bench.py --steps 200 --no-cpu-baseline --e2e-steps 1 > gpurun_out/b.json 2> gpurun_out/b.err
  echo "chunk $ch: $(python -c "import json;d=json.load(open('gpurun_out/b.json'));print(d['ms_per_step'])")"
done
