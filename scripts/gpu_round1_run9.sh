mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest9.log
timeout 900 python bench.py > gpurun_out/bench9.json 2> gpurun_out/bench9.err
timeout 600 python bench.py --impl reference --steps 16 > gpurun_out/bench9_ref.json 2> gpurun_out/bench9_ref.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 540 --launch-count 9 -o gpurun_out/prof_frame_r1f python bench.py --steps 3 --warmup 1 --preroll 60 --e2e-steps 4 --cpu-steps 0 --harness-frames 0 > gpurun_out/ncu_frame9.log 2>&1
python -c "
import json
j=json.loads(open('gpurun_out/bench9.json').read().strip().splitlines()[-1])
print('fps=%.0f ms=%.3f e2e=%.0f'%(j['value'],j['ms_per_step'],j['e2e']['value']), {k:round(v*1000) for k,v in j['stage_ms'].items()}, 'int_us=%.1f frac=%.3f'%(j['roofline']['mean_launch_us'], j['roofline']['frac']))
print(json.dumps(j['itmlib_harness'])[:700]); print(j['cpu_baseline'])
"
cat gpurun_out/bench9_ref.json | cut -c1-300
