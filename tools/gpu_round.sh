#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "widths or u32 or mirror or golden or compaction" > gpurun_out/pytest_widths.log 2>&1; tail -4 gpurun_out/pytest_widths.log | cut -c1-250
