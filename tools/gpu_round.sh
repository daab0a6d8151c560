#!/bin/bash
# Dev helper run under gpurun (edit per experiment).  Output -> gpurun_out/
mkdir -p gpurun_out
timeout 120 python bench.py --steps 100 --e2e-steps 2 > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err
python -c "import json;d=json.load(open('gpurun_out/bench_last.json'));print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'oracle sample', d['config']['results_match_oracle_sample'], 'cpu', d.get('cpu_baseline',{}).get('value'))"
