#!/bin/bash
# end-of-round verification on one B200: full GPU test suite, smoke, both bench arms, ncu launch list of the bench command
mkdir -p gpurun_out
T=${RF_TAG:-final}
t0=$(date +%s)
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/pytest_$T.log
python __graft_entry__.py smoke 2>&1 | tail -2
t0=$(date +%s)
python bench.py --impl reference > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err; echo "reference arm rc=$? ($(( $(date +%s) - t0 )) s)"; cut -c1-300 gpurun_out/bench_ref_$T.json
t0=$(date +%s)
python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"; cut -c1-400 gpurun_out/bench_$T.json
if [ -z "$RF_SKIP_NCU" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu_$T.log 2>&1
  echo "ncu launch list rc=$?"; wc -l gpurun_out/launches_$T.csv
fi
