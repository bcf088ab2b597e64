#!/bin/bash
# round 2: stand-alone CreateExpectedDepths (fast form) parity + shim timing, tail depth of the integrate kernel
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_interfaces.py tests/test_gpu_itm_harness.py -q -x 2>&1 | tail -15 > gpurun_out/pytest_parity8.log
tail -3 gpurun_out/pytest_parity8.log
for v in 3 2 1; do
  B200_V4_TAIL=$v timeout 300 python scripts/probe_trace.py > gpurun_out/trace_8_tail$v.txt 2>&1
  echo "== tail $v"; grep "^k_integrate" gpurun_out/trace_8_tail$v.txt | tail -4
done
timeout 600 python scripts/harness_repeat.py > gpurun_out/harness_8.json 2> gpurun_out/harness_8.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/harness_8.json').read().strip().splitlines()[-1])
    for k in ('reference_cuda_build','b200_itm_shim'):
        print(k, d[k]['median_fps'], json.dumps(d[k]['stages_us']))
except Exception as ex:
    print("harness unreadable", ex)
PY
