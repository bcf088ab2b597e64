#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_interfaces.py -q -x 2>&1 | tail -15 > gpurun_out/pytest_parity3.log
tail -3 gpurun_out/pytest_parity3.log
timeout 300 python scripts/probe_trace.py > gpurun_out/trace_3.txt 2>&1
grep -A10 "frame 5" gpurun_out/trace_3.txt
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_itm_harness.py tests/test_gpu_configs.py -q 2>&1 | tail -25 > gpurun_out/pytest_rest3.log
tail -4 gpurun_out/pytest_rest3.log
timeout 900 python bench.py > gpurun_out/bench_3.json 2> gpurun_out/bench_3.err
tail -2 gpurun_out/bench_3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_3.json').read().strip().splitlines()[-1])
for k in ['value','ms_per_step','parity_checked','parity','stage_ms','roofline','roofline_hires','decay_sweep','cpu_baseline','e2e','itmlib_harness']:
    print(k, json.dumps(d.get(k))[:700])
PY
