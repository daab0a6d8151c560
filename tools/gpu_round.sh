#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py -x -q -k "packed or len8 or typed or u32_streaming" > gpurun_out/pytest_r2k.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_r2k.log
timeout 600 python bench.py --steps 20 --warmup 5 --configs "" > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2k.json') if l.startswith('{')][-1])
e=d['e2e']; print('value',d['value'],'e2e packed6',e['value'],e['ms_per_step'],e['h2d_gbs'],'pack s',e['one_time_host_pack6_s'],'len8',e['len8']['value'],'csr',e['csr_u32']['value'],d['run'])
P
tail -3 gpurun_out/bench_r2k.err
