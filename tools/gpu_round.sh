#!/bin/bash
# Dev helper run under gpurun: microbench, GPU parity tests, bench A/B.  Output -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,pcie.link.gen.max,pcie.link.width.max --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
./tools/ubench/pipes > gpurun_out/pipes.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_dp4a.json 2> gpurun_out/bench_dp4a.err
RF_LIB_PATH=$PWD/rapidfuzz-rs_b200/lib/librfgpu_nodp4a.so timeout 300 python bench.py --steps 100 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_nodp4a.json 2> gpurun_out/bench_nodp4a.err
cat gpurun_out/pipes.txt
cut -c1-1500 gpurun_out/bench_dp4a.json
cut -c1-400 gpurun_out/bench_nodp4a.json
