#!/bin/bash
# Dev helper run under gpurun (edit per experiment).  Output -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1
nproc > gpurun_out/host_n8.txt; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/host_n8.txt; free -g | head -2 >> gpurun_out/host_n8.txt
python tools/pcie_peak.py --gpus 8 > gpurun_out/pcie_peak_n8.json 2> gpurun_out/pcie_peak_n8.err; echo "pcie rc=$?"; cut -c1-600 gpurun_out/pcie_peak_n8.json
timeout 900 python bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2e_n8.json 2> gpurun_out/bench_r2e_n8.err; echo "bench n8 rc=$?"
RF_BENCH_NUMA=0 timeout 600 python bench.py --gpus 8 --steps 20 --warmup 5 --configs "" --no-cpu-baseline > gpurun_out/bench_r2e_n8_nonuma.json 2> gpurun_out/bench_r2e_n8_nonuma.err; echo "bench n8 nonuma rc=$?"
python - <<'P'
import json
for f in ('gpurun_out/bench_r2e_n8.json','gpurun_out/bench_r2e_n8_nonuma.json'):
    try:
        txt=[l for l in open(f) if l.startswith('{')][-1]
        d=json.loads(txt)
        print(f,'value',d['value'],'e2e',d['e2e']['value'],'csr',d['e2e']['csr_u32']['value'],d['run'])
        for k,v in d['configs'].items(): print(k, {a:b for a,b in v.items() if a not in ('workload','roofline','what','collective')})
    except Exception as e: print(f, 'ERR', e)
P
tail -3 gpurun_out/bench_r2e_n8.err
python -m pytest tests/test_gpu_round2.py -x -q -k "sharded" > gpurun_out/pytest_r2e_n8.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2e_n8.log
