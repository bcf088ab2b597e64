#!/bin/bash
# round 2, second GPU session: parity of the new allocation / decay kernels, launch trace, full-frame ncu capture with source counters
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -15 > gpurun_out/pytest_parity.log
tail -4 gpurun_out/pytest_parity.log
timeout 300 python scripts/probe_trace.py > gpurun_out/trace_2.txt 2>&1
grep -A12 "frame 5" gpurun_out/trace_2.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_all.log
tail -6 gpurun_out/pytest_all.log
COMMON="--preroll 215 --steps 3 --warmup 1 --e2e-steps 4 --e2e-raw-steps 0 --cpu-steps 0 --harness-frames 0 --hires-frames 0 --profile-step 1 --no-parity-check"
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_ \
   -o gpurun_out/prof_frame_r2a python bench.py $COMMON > gpurun_out/ncu_frame_r2a.log 2>&1
ls -la gpurun_out/*.ncu-rep
