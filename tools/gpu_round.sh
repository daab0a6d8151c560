#!/bin/bash
# Dev helper run under gpurun (edit per experiment).  Output -> gpurun_out/
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py -x -q -k "sharded or comm or u32_streaming or typed" > gpurun_out/pytest_r2h_n8.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_r2h_n8.log
timeout 900 python tools/bench_sharded_abi.py --gpus 8 --per-gpu 50000000 > gpurun_out/sharded_abi_n8.json 2> gpurun_out/sharded_abi_n8.err; echo "abi rc=$?"
tail -3 gpurun_out/sharded_abi_n8.err
timeout 900 python bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2h_n8.json 2> gpurun_out/bench_r2h_n8.err; echo "bench n8 rc=$?"
python - <<'P'
import json
for f in ('gpurun_out/sharded_abi_n8.json','gpurun_out/bench_r2h_n8.json'):
    try:
        txt=[l for l in open(f) if l.startswith('{')][-1]
        d=json.loads(txt)
        if 'configs' in d:
            print('value',d['value'],'e2e',d['e2e']['value'],'csr',d['e2e']['csr_u32']['value'])
            print(json.dumps(d['configs']['c2_gather'],indent=1)[:1800])
            print({k:v for k,v in d['configs']['c5'].items() if k not in ('workload','roofline','collective')})
        else:
            print(json.dumps(d,indent=1)[:3000])
    except Exception as e: print(f,'ERR',e)
P
tail -3 gpurun_out/bench_r2h_n8.err
