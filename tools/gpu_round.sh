#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2s_n8.json 2> gpurun_out/bench_r2s_n8.err; echo "bench n8 rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2s_n8.json') if l.startswith('{')][-1])
e=d['e2e']; print('value',d['value'],'e2e',e['value'],e['ms_per_step'],e['h2d_gbs'],'len8',e['len8']['value'],'csr',e['csr_u32']['value'],d['run'], 'pack', e['one_time_host_pack6_s'])
for k,v in d['configs'].items(): print(k, {a:b for a,b in v.items() if a in ('ms_per_step','pairs_per_s','error','skipped','scan_ms_max_over_ranks','gather_merge_ms')})
print(d['configs']['c2_gather'].get('rf_batch_score_u32_allgather_device'))
P
tail -3 gpurun_out/bench_r2s_n8.err
timeout 600 python bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_r2s_n4.json 2> gpurun_out/bench_r2s_n4.err; echo "bench n4 rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2s_n4.json') if l.startswith('{')][-1])
e=d['e2e']; print('N4 value',d['value'],'e2e',e['value'],'len8',e['len8']['value'],'csr',e['csr_u32']['value'])
for k,v in d['configs'].items(): print(k, {a:b for a,b in v.items() if a in ('ms_per_step','pairs_per_s','error','skipped')})
P
