#!/bin/bash
# Dev helper run under gpurun (edit per experiment).  Output -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n2.txt 2>&1
python -m pytest tests/test_gpu_round2.py -x -q -k "sharded or register_column" > gpurun_out/pytest_r2d.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_r2d.log
python tools/bench_configs.py mw 2>gpurun_out/r2_mw_new3.err | head -3 > gpurun_out/r2_mw_new3.jsonl; cut -c1-330 gpurun_out/r2_mw_new3.jsonl
timeout 900 python bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2d_n2.json 2> gpurun_out/bench_r2d_n2.err; echo "bench n2 rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_r2d_n2.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'csr',d['e2e']['csr_u32']['value'],d['run'])
for k,v in d['configs'].items(): print(k, {a:b for a,b in v.items() if a not in ('workload','roofline','what','collective')})
P
tail -3 gpurun_out/bench_r2d_n2.err
python tools/pcie_peak.py --gpus 2 > gpurun_out/pcie_peak_n2.json 2> gpurun_out/pcie_peak_n2.err; cut -c1-1500 gpurun_out/pcie_peak_n2.json
