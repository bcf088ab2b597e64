#!/bin/bash
# round 2, final evidence session: the whole GPU suite, the default bench, one ncu frame capture
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu_final.log
tail -5 gpurun_out/pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -2 gpurun_out/bench_final.err
COMMON="--preroll 215 --steps 3 --warmup 1 --e2e-steps 4 --e2e-raw-steps 0 --cpu-steps 0 --harness-frames 0 --hires-frames 0 --decay-blocks 0 --profile-step 1 --no-parity-check"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_ \
   -o gpurun_out/prof_frame_r2c python bench.py $COMMON > gpurun_out/ncu_frame_r2c.log 2>&1
timeout 300 python scripts/probe_trace.py > gpurun_out/trace_final.txt 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
for k in ['value','ms_per_step','parity_checked','stage_ms','roofline','roofline_hires','decay_sweep','cpu_baseline','e2e','e2e_raw','itmlib_harness','view_builder','meshing','frames_ops']:
    print(k, json.dumps(d.get(k))[:420])
PY
