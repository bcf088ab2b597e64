mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_view.py tests/test_gpu_frames.py tests/test_gpu_itm_harness.py -m gpu -q -s 2>&1 | tail -12 | tee gpurun_out/pytest16.log
timeout 900 python bench.py --steps 200 --cpu-steps 0 --harness-frames 40 --hires-frames 0 > gpurun_out/bench16.json 2> gpurun_out/bench16.err
tail -3 gpurun_out/bench16.err
python -c "
import json
j=json.loads(open('gpurun_out/bench16.json').read().strip().splitlines()[-1])
print('fps=%.0f ms=%.3f e2e=%.0f'%(j['value'],j['ms_per_step'],j['e2e']['value']), {k:round(v*1000) for k,v in j['stage_ms'].items()})
print('e2e_raw', j.get('e2e_raw'))
vb=j.get('view_builder') or {}
print('view_builder', {k:v for k,v in vb.items() if k!='what'})
print('frames_ops', {k:({kk:vv for kk,vv in v.items() if kk!='what'} if isinstance(v,dict) else v) for k,v in (j.get('frames_ops') or {}).items()})
"
