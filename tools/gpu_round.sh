#!/bin/bash
# Dev helper run under gpurun.  Output -> gpurun_out/
mkdir -p gpurun_out
timeout 900 python tools/bench_configs.py osa,simple > gpurun_out/cfg_osa_simple.jsonl 2> gpurun_out/cfg_osa_simple.err
python - <<'PY'
import json
for l in open('gpurun_out/cfg_osa_simple.jsonl'):
    d=json.loads(l); print(d['config'], round(d['ms_per_step'],3), 'ms', '%.3g'%d['pairs_per_s'], d.get('matches_oracle_sample'))
PY
tail -2 gpurun_out/cfg_osa_simple.err
