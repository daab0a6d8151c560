#!/bin/bash
# dev helper: one gpurun call = targeted parity tests + the shared-corpus bench of the kernels touched last
mkdir -p gpurun_out
T=${RF_TAG:-r2x}
python -m pytest tests -m gpu -x -q -k "${RF_K:-jaro or single_word or options_are or golden}" > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$T.log
for o in ${RF_OPTS_LIST:-none}; do
  if [ "$o" = none ]; then o=""; fi
  RF_OPTS=$o python tools/bench_shared_corpus.py ${RF_CASES:-jw32,jw48,jaro64} 2>> gpurun_out/shared_$T.err | tee -a gpurun_out/shared_$T.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['process_options'], d['case'], round(d['ms_per_step'], 4), d['bit_exact_vs_oracle_sample'])"
done
