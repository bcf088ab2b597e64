#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_interfaces.py tests/test_gpu_mesh.py tests/test_gpu_view.py -q -x 2>&1 | tail -15 > gpurun_out/pytest_parity4.log
tail -3 gpurun_out/pytest_parity4.log
timeout 300 python scripts/probe_trace.py > gpurun_out/trace_4.txt 2>&1
grep -A12 "frame 5" gpurun_out/trace_4.txt
grep -A8 "UpdateView" gpurun_out/trace_4.txt
B200_V4_PPL=2 timeout 300 python scripts/probe_trace.py > gpurun_out/trace_4_ppl2.txt 2>&1
grep -A12 "frame 5" gpurun_out/trace_4_ppl2.txt | grep integrate
B200_INTEGRATE=v3 timeout 300 python scripts/probe_trace.py > gpurun_out/trace_4_v3.txt 2>&1
grep -A12 "frame 5" gpurun_out/trace_4_v3.txt | grep integrate
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_itm_harness.py tests/test_gpu_configs.py -q 2>&1 | tail -25 > gpurun_out/pytest_rest4.log
tail -4 gpurun_out/pytest_rest4.log
timeout 900 python bench.py > gpurun_out/bench_4.json 2> gpurun_out/bench_4.err
tail -2 gpurun_out/bench_4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_4.json').read().strip().splitlines()[-1])
for k in ['value','ms_per_step','parity_checked','stage_ms','roofline','roofline_hires','decay_sweep','cpu_baseline','e2e','itmlib_harness']:
    print(k, json.dumps(d.get(k))[:600])
PY
