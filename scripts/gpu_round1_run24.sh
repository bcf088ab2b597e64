mkdir -p gpurun_out
B200_GRAPH=1 timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest24.log
for g in 1 0; do
B200_GRAPH=$g timeout 900 python bench.py --steps 300 --cpu-steps 0 --harness-frames 0 --hires-frames 0 > gpurun_out/bench24_$g.json 2> gpurun_out/bench24_$g.err
tail -2 gpurun_out/bench24_$g.err
python -c "
import json
j=json.loads(open('gpurun_out/bench24_$g.json').read().strip().splitlines()[-1])
print('graph $g: fps=%.0f ms=%.3f e2e=%.0f raw=%.0f int_us=%.1f rays=%d'%(j['value'],j['ms_per_step'],j['e2e']['value'],j['e2e_raw']['value'],j['roofline']['mean_launch_us'],j['rays_hit']), {k:round(v*1000) for k,v in j['stage_ms'].items()})
"
done
