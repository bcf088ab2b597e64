mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "pipelined or fused" 2>&1 | tail -5 | tee gpurun_out/pytest13.log
timeout 900 python bench.py --steps 300 --cpu-steps 0 --harness-frames 0 --hires-frames 0 --e2e-steps 203 > gpurun_out/bench13.json 2> gpurun_out/bench13.err
python -c "
import json
j=json.loads(open('gpurun_out/bench13.json').read().strip().splitlines()[-1])
print('fps=%.0f ms=%.3f e2e=%.0f'%(j['value'],j['ms_per_step'],j['e2e']['value']), {k:round(v*1000) for k,v in j['stage_ms'].items()})
"
