mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest10.log
timeout 900 python bench.py --cpu-steps 0 --harness-frames 0 > gpurun_out/bench10.json 2> gpurun_out/bench10.err
B200_NO_OVERLAP=1 timeout 900 python bench.py --cpu-steps 0 --harness-frames 0 --hires-frames 0 > gpurun_out/bench10_nooverlap.json 2> gpurun_out/bench10_nooverlap.err
python -c "
import json
for f in ['gpurun_out/bench10.json','gpurun_out/bench10_nooverlap.json']:
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(f,'fps=%.0f ms=%.3f e2e=%.0f'%(j['value'],j['ms_per_step'],j['e2e']['value']), {k:round(v*1000) for k,v in j['stage_ms'].items()}, 'int_us=%.1f frac=%.3f'%(j['roofline']['mean_launch_us'], j['roofline']['frac']))
    print(j.get('roofline_hires'))
"
tail -3 gpurun_out/bench10.err
