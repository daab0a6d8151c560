#!/bin/bash
# dev helper: one gpurun call = targeted parity tests + the shared-corpus bench of the kernels touched last
mkdir -p gpurun_out
T=${RF_TAG:-r2x}
python -m pytest tests -m gpu -x -q -k "jaro or single_word or hamming or band or config3 or multi_word or register_column or options_are" > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$T.log
python tools/bench_shared_corpus.py ${RF_CASES:-ham,pre,post,jw32,jw48,jw48off,jaro64,lev32,indel32} 2> gpurun_out/shared_$T.err | tee gpurun_out/shared_$T.jsonl | cut -c1-260
python tools/bench_configs.py c3 2>/dev/null | tee gpurun_out/c3_$T.jsonl | cut -c1-300
