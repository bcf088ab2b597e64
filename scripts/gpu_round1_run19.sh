mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest19.log
timeout 600 python scripts/probe_trace.py 2>&1 | grep -A12 "frame 5" | head -13 | tee gpurun_out/probe19.log
for v in "56 0" "48 0" "48 1"; do
  set -- $v
  B200_INTEGRATE_REGS=$1 B200_GRAPH=$2 timeout 900 python bench.py --steps 300 --cpu-steps 0 --harness-frames 0 --hires-frames 0 --e2e-raw-steps 0 --e2e-steps 53 > gpurun_out/bench19_$1_$2.json 2> gpurun_out/bench19_$1_$2.err
  python -c "
import json
j=json.loads(open('gpurun_out/bench19_$1_$2.json').read().strip().splitlines()[-1])
print('regs $1 graph $2: fps=%.0f ms=%.3f e2e=%.0f int_us=%.1f'%(j['value'],j['ms_per_step'],j['e2e']['value'],j['roofline']['mean_launch_us']), {k:round(v*1000) for k,v in j['stage_ms'].items()})
"
done
