#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_interfaces.py tests/test_gpu_eval.py -q -x 2>&1 | tail -15 > gpurun_out/pytest_parity5.log
tail -3 gpurun_out/pytest_parity5.log
timeout 300 python scripts/probe_trace.py > gpurun_out/trace_5.txt 2>&1
grep -A24 "frame 5" gpurun_out/trace_5.txt | cut -c1-900
