#!/bin/bash
mkdir -p gpurun_out
for path in 0 2 0 2; do
  RF_W1_PATH=$path timeout 300 python bench.py --steps 200 --no-cpu-baseline --e2e-steps 1 > gpurun_out/b$path.json 2> gpurun_out/b.err
  echo "path $path: $(python -c "import json;d=json.load(open('gpurun_out/b$path.json'));print(d['ms_per_step'], d['config']['results_match_oracle_sample'])")"; tail -2 gpurun_out/b.err
done
