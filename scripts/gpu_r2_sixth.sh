#!/bin/bash
# round 2, evidence session: parity, launch trace, ncu captures (one frame with source counters; the 4 mm integrate launch;
# the launch list of the bench command), default bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_interfaces.py -q -x 2>&1 | tail -15 > gpurun_out/pytest_parity6.log
tail -3 gpurun_out/pytest_parity6.log
timeout 300 python scripts/probe_trace.py > gpurun_out/trace_6.txt 2>&1
grep -A22 "frame 5" gpurun_out/trace_6.txt | cut -c1-260
COMMON="--preroll 215 --steps 3 --warmup 1 --e2e-steps 4 --e2e-raw-steps 0 --cpu-steps 0 --harness-frames 0 --hires-frames 0 --decay-blocks 0 --profile-step 1 --no-parity-check"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_ \
   -o gpurun_out/prof_frame_r2b python bench.py $COMMON > gpurun_out/ncu_frame_r2b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_integrate_v4 --launch-skip 10 --launch-count 1 \
   -o gpurun_out/prof_integrate_hires_r2b python scripts/profile_hires.py 12 > gpurun_out/ncu_hires_r2b.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2b.csv \
   python bench.py --steps 20 --warmup 3 --preroll 215 --e2e-steps 4 --e2e-raw-steps 0 --cpu-steps 0 --harness-frames 0 --hires-frames 0 --decay-blocks 0 --no-parity-check > gpurun_out/launches_r2b.log 2>&1
timeout 1200 python bench.py > gpurun_out/bench_6.json 2> gpurun_out/bench_6.err
tail -2 gpurun_out/bench_6.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_6.json').read().strip().splitlines()[-1])
for k in ['value','ms_per_step','parity_checked','stage_ms','roofline','roofline_hires','decay_sweep','cpu_baseline','e2e','itmlib_harness','view_builder']:
    print(k, json.dumps(d.get(k))[:500])
PY
