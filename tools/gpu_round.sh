#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py -x -q -k "sharded_stream" > gpurun_out/pytest_r2v.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2v.log
timeout 900 python tools/bench_sharded_abi.py --gpus 8 --per-gpu 50000000 --skip-c5 --skip-gather > gpurun_out/sharded_abi_n8_dyn3.json 2> gpurun_out/sharded_abi_n8_dyn3.err; echo "abi rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/sharded_abi_n8_dyn3.json') if l.startswith('{')][-1])
for k in ('stream_from_pinned_host','stream_len8_dynamic','stream_len8_packed6_dynamic'): print(k, d.get(k))
P
tail -3 gpurun_out/sharded_abi_n8_dyn3.err
