mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest18.log
timeout 600 python scripts/probe_trace.py 2>&1 | grep -A12 "frame 5" | tee gpurun_out/probe18.log
timeout 900 python bench.py --steps 300 --cpu-steps 0 --harness-frames 0 --hires-frames 0 --e2e-raw-steps 0 > gpurun_out/bench18.json 2> gpurun_out/bench18.err
python -c "
import json
j=json.loads(open('gpurun_out/bench18.json').read().strip().splitlines()[-1])
print('fps=%.0f ms=%.3f e2e=%.0f'%(j['value'],j['ms_per_step'],j['e2e']['value']), {k:round(v*1000) for k,v in j['stage_ms'].items()})
"
