mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest7.log
timeout 600 python bench.py --steps 400 --cpu-steps 0 --harness-frames 0 > gpurun_out/bench7_tma.json 2> gpurun_out/bench7_tma.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 600 --launch-count 10 -o gpurun_out/prof_frame_r1e python bench.py --steps 3 --warmup 1 --preroll 60 --e2e-steps 2 --cpu-steps 0 --harness-frames 0 > gpurun_out/ncu_frame7.log 2>&1
python -c "
import json
j=json.loads(open('gpurun_out/bench7_tma.json').read().strip().splitlines()[-1])
print('fps=%.0f ms=%.3f e2e=%.0f'%(j['value'],j['ms_per_step'],j['e2e']['value']), {k:round(v*1000) for k,v in j['stage_ms'].items()}, 'int_us=%.1f'%j['roofline']['mean_launch_us'])
"
