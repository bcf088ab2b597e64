# dynamic in-flight depth in the V3 producer: parity (three sequences) + one short bench
mkdir -p gpurun_out
(timeout 100 python -m pytest tests/test_gpu_parity.py -q -x -k "kitti_sequence or integrate_wide or depth_weighting or hash_collisions" 2>&1 | tail -4 > gpurun_out/pytest35.log; cat gpurun_out/pytest35.log) &
timeout 120 python bench.py --steps 100 --preroll 205 --cpu-steps 0 --harness-frames 0 --e2e-steps 8 --e2e-raw-steps 0 --hires-frames 16 > gpurun_out/bench35.json 2> gpurun_out/bench35.err
wait
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench35.json').read().strip().splitlines()[-1])
print('fps=%.0f ms=%.3f int_us=%.1f frac=%.3f'%(j['value'],j['ms_per_step'],j['roofline']['mean_launch_us'],j['roofline']['frac']), {k:round(v*1000) for k,v in j['stage_ms'].items()})
print('hires', {k:v for k,v in (j.get('roofline_hires') or {}).items() if k in ('mean_launch_us','frac','error')})
PY
