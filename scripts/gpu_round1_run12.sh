mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/pytest12.log
for t in 3 4 5 2; do
  B200_RC_TILE=$t timeout 600 python bench.py --steps 200 --cpu-steps 0 --harness-frames 0 --hires-frames 0 --e2e-steps 8 > gpurun_out/bench12_tile$t.json 2> gpurun_out/bench12_tile$t.err
  python -c "
import json
j=json.loads(open('gpurun_out/bench12_tile$t.json').read().strip().splitlines()[-1])
print('tile $t fps=%.0f ms=%.3f'%(j['value'],j['ms_per_step']), {k:round(v*1000) for k,v in j['stage_ms'].items()})
"
done
