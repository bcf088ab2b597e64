# V3 integrate with one-block-ahead projection + L2 image prefetch: parity, bench for 2 / 3 / 4 resident CTAs.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "integrate or kitti_sequence or depth_weighting or config1 or fused_async or only_update" 2>&1 | tail -15 | tee gpurun_out/pytest31.log
B="--steps 150 --preroll 215 --cpu-steps 0 --harness-frames 0 --e2e-steps 23 --e2e-raw-steps 0 --hires-frames 24"
for v in "v3 3" "v3 2" "v3 4"; do
  set -- $v
  B200_INTEGRATE_IMPL=$1 B200_V3_CTAS=$2 timeout 600 python bench.py $B > gpurun_out/bench31_$1_$2.json 2> gpurun_out/bench31_$1_$2.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench31_$1_$2.json').read().strip().splitlines()[-1])
    print('$1 ctas$2: fps=%.0f ms=%.3f e2e=%.0f int_us=%.1f frac=%.3f'%(j['value'],j['ms_per_step'],j['e2e']['value'],j['roofline']['mean_launch_us'],j['roofline']['frac']), {k:round(v*1000) for k,v in j['stage_ms'].items()})
    print('   hires', {k:v for k,v in (j.get('roofline_hires') or {}).items() if k in ('mean_launch_us','achieved','frac','visible_blocks','error')}, j['clocks'])
except Exception as ex:
    print('$1 $2: failed', ex); print(open('gpurun_out/bench31_$1_$2.err').read()[-1500:])
PY
done
