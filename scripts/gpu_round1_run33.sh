# Round-1 evidence run: full GPU test-suite, default bench, ncu launch list, full ncu capture of one steady-state frame.
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest33.log
timeout 400 python bench.py > gpurun_out/bench33.json 2> gpurun_out/bench33.err
tail -2 gpurun_out/bench33.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench33.json').read().strip().splitlines()[-1])
print('fps=%.0f ms=%.3f e2e=%.0f raw=%.0f int_us=%.1f frac=%.3f'%(j['value'],j['ms_per_step'],j['e2e']['value'],j['e2e_raw']['value'],j['roofline']['mean_launch_us'],j['roofline']['frac']), {k:round(v*1000) for k,v in j['stage_ms'].items()})
print('harness', {k:(round(v['value']) if isinstance(v,dict) else v) for k,v in (j.get('itmlib_harness') or {}).items() if k!='what'})
print('cpu', j['cpu_baseline']['value'], j['cpu_baseline']['cores'])
print('hires', {k:v for k,v in (j.get('roofline_hires') or {}).items() if k in ('mean_launch_us','achieved','frac','visible_blocks')})
print('clocks', j['clocks'])
PY
COMMON="--preroll 215 --steps 3 --warmup 1 --e2e-steps 4 --e2e-raw-steps 0 --cpu-steps 0 --harness-frames 0 --hires-frames 0"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --launch-skip 2190 --launch-count 84 --csv \
   --log-file gpurun_out/r01h_launches.csv python bench.py $COMMON > gpurun_out/ncu_launches.log 2>&1
timeout 420 ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 2202 --launch-count 13 \
   -o gpurun_out/prof_frame_r1h python bench.py $COMMON > gpurun_out/ncu_frame.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r01h_launches.csv
