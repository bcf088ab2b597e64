#!/bin/bash
# round 2, first GPU session: pipe throughputs, fresh-process ITMLib harness repeats, launch trace of the fused frame, baseline tests + bench
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_first.txt 2>&1
timeout 120 scripts/ubench/pipes.bin > gpurun_out/pipes.txt 2>&1
timeout 300 python scripts/harness_repeat.py 330 230 3 > gpurun_out/harness_repeat.json 2> gpurun_out/harness_repeat.err
timeout 300 python scripts/harness_repeat.py 160 60 3 > gpurun_out/harness_repeat_short.json 2>> gpurun_out/harness_repeat.err
timeout 300 python scripts/probe_trace.py > gpurun_out/trace_base.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_base.log
timeout 600 python bench.py > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err
cat gpurun_out/pipes.txt; cat gpurun_out/harness_repeat.json; tail -3 gpurun_out/pytest_base.log
