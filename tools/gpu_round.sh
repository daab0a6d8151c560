#!/bin/bash
# dev helper: one gpurun call = targeted parity tests + the shared-corpus bench of the kernels touched last
mkdir -p gpurun_out
T=${RF_TAG:-r2x}
python -m pytest tests -m gpu -x -q -k "${RF_K:-jaro or single_word or options_are or golden}" > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$T.log
python tools/bench_shared_corpus.py ${RF_CASES:-jw32,jw32pair,jw32r64,jw48,jw48pair,jaro64} 2> gpurun_out/shared_$T.err | tee gpurun_out/shared_$T.jsonl | cut -c1-200
