#!/bin/bash
# dev helper: one `ncu --set full` capture per case of tools/bench_shared_corpus.py (2e7-candidate corpus); the reports are
# summarised on the box (tools/ncu_summary.py) and only the text comes back (gpurun_out is capped at 64 MiB)
mkdir -p gpurun_out /tmp/ncu
export RF_CFG_SCALE=${RF_CFG_SCALE:-0.2}
for spec in ${RF_NCU:-indel32:scan_lb_kernel ham:hamming_lb_kernel jw48:scan_jaro64_kernel jw32:scan_jaro32_kernel}; do
  c=${spec%%:*}; k=${spec##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o /tmp/ncu/$c \
    python tools/bench_shared_corpus.py $c > /tmp/ncu/$c.log 2>&1
  echo "$c rc=$?"
  { echo "ncu --set full --clock-control none, launch 4 of: RF_CFG_SCALE=$RF_CFG_SCALE python tools/bench_shared_corpus.py $c  (kernel regex $k)"
    python tools/ncu_summary.py /tmp/ncu/$c.ncu-rep; } > gpurun_out/ncu_${RF_TAG:-r2}_${c}_summary.txt 2>&1
  head -12 gpurun_out/ncu_${RF_TAG:-r2}_${c}_summary.txt
done
