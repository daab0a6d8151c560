#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/bench_configs.py simple > gpurun_out/cfg_simple.jsonl 2> gpurun_out/cfg_simple.err
cat gpurun_out/cfg_simple.jsonl; tail -3 gpurun_out/cfg_simple.err
