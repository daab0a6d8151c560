#!/bin/bash
# Dev helper run under gpurun (edit per experiment).  Output -> gpurun_out/
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py -x -q > gpurun_out/pytest_r2g.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_r2g.log
timeout 900 python tools/bench_sharded_abi.py --gpus 2 --per-gpu 100000000 > gpurun_out/sharded_abi_n2.json 2> gpurun_out/sharded_abi_n2.err; echo "abi rc=$?"
python - <<'P'
import json
txt=[l for l in open('gpurun_out/sharded_abi_n2.json') if l.startswith('{')][-1]
print(json.dumps(json.loads(txt),indent=1))
P
tail -5 gpurun_out/sharded_abi_n2.err
