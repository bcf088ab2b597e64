#!/bin/bash
# Short A/B of the IntegrateIntoScene variants and knobs on one box: parity subset first, then one short bench per setting.
#   B200_INTEGRATE_IMPL = v3 (default) | tma | ldg      B200_V3_CTAS = 3 (default) | 2 | 4      B200_V3_PREFETCH = 1 | 0
#   B200_RC_ORDER = 1 (default, principal-point rows first) | 0
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "division or integrate or kitti_sequence or depth_weighting or config1 or fused_async" 2>&1 | tail -6
B="--steps 150 --preroll 215 --cpu-steps 0 --harness-frames 0 --e2e-steps 23 --e2e-raw-steps 0 --hires-frames 24"
for v in "v3 3" "v3 4" "v3 2" "tma 4"; do
  set -- $v
  B200_INTEGRATE_IMPL=$1 B200_V3_CTAS=$2 timeout 600 python bench.py $B > gpurun_out/cmp_$1_$2.json 2> gpurun_out/cmp_$1_$2.err
  python - <<PY
import json
j=json.loads(open('gpurun_out/cmp_$1_$2.json').read().strip().splitlines()[-1])
print('$1 ctas$2: fps=%.0f e2e=%.0f int_us=%.1f frac=%.3f'%(j['value'],j['e2e']['value'],j['roofline']['mean_launch_us'],j['roofline']['frac']), {k:round(v*1000) for k,v in j['stage_ms'].items()})
print('   hires', {k:v for k,v in (j.get('roofline_hires') or {}).items() if k in ('mean_launch_us','achieved','frac')})
PY
done
