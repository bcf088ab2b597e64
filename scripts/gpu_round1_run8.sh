mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest8.log
timeout 900 python bench.py > gpurun_out/bench8.json 2> gpurun_out/bench8.err
timeout 600 python bench.py --impl reference --steps 16 > gpurun_out/bench8_ref.json 2> gpurun_out/bench8_ref.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke8.log 2>&1
tail -2 gpurun_out/smoke8.log
python -c "
import json
j=json.loads(open('gpurun_out/bench8.json').read().strip().splitlines()[-1])
print('fps=%.0f ms=%.3f e2e=%.0f'%(j['value'],j['ms_per_step'],j['e2e']['value']), {k:round(v*1000) for k,v in j['stage_ms'].items()}, 'int_us=%.1f frac=%.3f'%(j['roofline']['mean_launch_us'], j['roofline']['frac']))
print(json.dumps(j['itmlib_harness'])[:700]); print(j['cpu_baseline'])
"
cat gpurun_out/bench8_ref.json | cut -c1-400
