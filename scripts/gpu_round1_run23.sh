mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest23.log
timeout 600 python scripts/probe_trace.py 2>&1 | grep -A12 "frame 5" | head -13 | tee gpurun_out/probe23.log
for g in 0 1; do
B200_GRAPH=$g timeout 900 python bench.py --steps 300 --cpu-steps 0 --harness-frames 0 --hires-frames 0 > gpurun_out/bench23_$g.json 2> gpurun_out/bench23_$g.err
python -c "
import json
j=json.loads(open('gpurun_out/bench23_$g.json').read().strip().splitlines()[-1])
print('graph $g: fps=%.0f ms=%.3f e2e=%.0f raw=%.0f int_us=%.1f'%(j['value'],j['ms_per_step'],j['e2e']['value'],j['e2e_raw']['value'],j['roofline']['mean_launch_us']), {k:round(v*1000) for k,v in j['stage_ms'].items()})
"
done
