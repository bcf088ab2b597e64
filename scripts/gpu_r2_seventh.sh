#!/bin/bash
# round 2: check of the marking / raycast micro-optimisations and the integrate variants (launch trace), then a short bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_interfaces.py -q -x 2>&1 | tail -15 > gpurun_out/pytest_parity7.log
tail -3 gpurun_out/pytest_parity7.log
for v in v4 fast v3; do
  B200_INTEGRATE_IMPL=$v timeout 300 python scripts/probe_trace.py > gpurun_out/trace_7_$v.txt 2>&1
  echo "== $v"; grep -A32 "frame 5" gpurun_out/trace_7_$v.txt | grep "^k_" | head -7
done
B200_V4_PPL=2 timeout 300 python scripts/probe_trace.py > gpurun_out/trace_7_ppl2.txt 2>&1
echo "== v4 ppl2"; grep -A32 "frame 5" gpurun_out/trace_7_ppl2.txt | grep "^k_integrate"
timeout 600 python bench.py --harness-frames 0 --hires-frames 0 --decay-blocks 0 --cpu-steps 0 > gpurun_out/bench_7.json 2> gpurun_out/bench_7.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_7.json').read().strip().splitlines()[-1])
for k in ['value','ms_per_step','parity_checked','stage_ms','e2e','meshing','frames_ops']:
    print(k, json.dumps(d.get(k))[:700])
PY
