#!/bin/bash
mkdir -p gpurun_out
for v in base evl; do
  if [ $v = evl ]; then export RF_LIB_PATH=$PWD/rapidfuzz-rs_b200/lib/librfgpu_evl.so; else unset RF_LIB_PATH; fi
  python bench.py --steps 50 --warmup 5 --configs "" --no-cpu-baseline --e2e-steps 1 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v lev ms', d['ms_per_step'], 'ok', d['run']['results_match_oracle_sample'])"
  python tools/bench_configs.py fam 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$v', d['config'], round(d['ms_per_step'],4), d['matches_oracle_sample'])"
  RF_CFG_SCALE=0.5 ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:scan_lb_kernel -s 3 -c 1 --csv python tools/bench_configs.py indel 2>/dev/null | grep -E "dram__bytes|gpu__time" | cut -d, -f13- | tr '\n' ' '; echo " <- $v ncu indel 0.5 scale"
done
