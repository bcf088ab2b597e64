#!/bin/bash
# One GPU-box session that produces everything the round's numbers come from (run through gpurun from the repo root):
#   1. the GPU test-suite, 2. the default bench line, 3. an `ncu --set full` capture of every kernel of ONE steady-state
#   frame (bench.py brackets timed step 1 with cudaProfilerStart/Stop), summarised by scripts/ncu_summary.py into profiles/.
# A number printed by a run under ncu is never a bench value.
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -2 gpurun_out/bench_$TAG.err
COMMON="--preroll 215 --steps 3 --warmup 1 --e2e-steps 4 --e2e-raw-steps 0 --cpu-steps 0 --harness-frames 0 --hires-frames 0 --profile-step 1"
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_ \
   -o gpurun_out/prof_frame_$TAG python bench.py $COMMON > gpurun_out/ncu_frame_$TAG.log 2>&1
# the 4 mm roofline-stress stream: one steady-state IntegrateIntoScene launch (13 launches build the map first)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_integrate --launch-skip 13 --launch-count 1 \
   -o gpurun_out/prof_integrate_hires_$TAG python scripts/profile_hires.py 16 > gpurun_out/ncu_hires_$TAG.log 2>&1
ls -la gpurun_out/*.ncu-rep
# back in the container:  python scripts/ncu_summary.py gpurun_out/prof_frame_$TAG.ncu-rep profiles/${TAG}_frame_kernels.md profiles/integrate_traffic.json
