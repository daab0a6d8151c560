#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_sharded_abi.py --gpus 8 --per-gpu 50000000 --only-gather > gpurun_out/sharded_abi_push_n8.json 2> gpurun_out/sharded_abi_push_n8.err; echo "abi rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/sharded_abi_push_n8.json') if l.startswith('{')][-1])
for k,v in d['score_allgather_device'].items(): print(k,v)
P
tail -3 gpurun_out/sharded_abi_push_n8.err
