mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest22.log
timeout 600 python scripts/probe_trace.py 2>&1 | grep -A12 "frame 5" | head -13 | tee gpurun_out/probe22.log
timeout 1200 python bench.py > gpurun_out/bench22.json 2> gpurun_out/bench22.err
tail -2 gpurun_out/bench22.err
python -c "
import json
j=json.loads(open('gpurun_out/bench22.json').read().strip().splitlines()[-1])
print('fps=%.0f ms=%.3f e2e=%.0f raw=%.0f int_us=%.1f frac=%.3f'%(j['value'],j['ms_per_step'],j['e2e']['value'],j['e2e_raw']['value'],j['roofline']['mean_launch_us'],j['roofline']['frac']), {k:round(v*1000) for k,v in j['stage_ms'].items()})
print('harness', {k:(round(v['value']) if isinstance(v,dict) else v) for k,v in (j.get('itmlib_harness') or {}).items() if k!='what'})
print('cpu', j['cpu_baseline']['value'], j['cpu_baseline']['cores'])
print('hires', {k:v for k,v in (j.get('roofline_hires') or {}).items() if k in ('mean_launch_us','achieved','frac','visible_blocks')})
vb=j.get('view_builder') or {}
print('view_builder', {k:v for k,v in vb.items() if k!='what'})
print('frames_ops', {k:({kk:round(vv,2) if isinstance(vv,float) else vv for kk,vv in v.items() if kk!='what'} if isinstance(v,dict) else v) for k,v in (j.get('frames_ops') or {}).items()})
print('clocks', j['clocks'])
"
