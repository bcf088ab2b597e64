mkdir -p gpurun_out
for d in a b c d e; do
timeout 900 python bench.py --steps 100 --preroll 210 --cpu-steps 0 --harness-frames 0 --hires-frames 0 > gpurun_out/bench27_$d.json 2> gpurun_out/bench27_$d.err
python -c "
import json
j=json.loads(open('gpurun_out/bench27_$d.json').read().strip().splitlines()[-1])
print('$d: fps=%.0f e2e=%.0f raw=%.0f'%(j['value'],j['e2e']['value'],j['e2e_raw']['value']), 'e2e max gap %.2f'%j['e2e']['frame_ms']['max'], 'raw max gap %.2f'%j['e2e_raw']['frame_ms']['max'], j['clocks'])
"
done
