mkdir -p gpurun_out
for d in none nosampler nogc none; do
B200_BENCH_DIAG=$d timeout 900 python bench.py --steps 100 --preroll 210 --cpu-steps 0 --harness-frames 0 --hires-frames 0 --e2e-steps 103 --e2e-raw-steps 103 > gpurun_out/bench25_$d.json 2> gpurun_out/bench25_$d.err
python -c "
import json
j=json.loads(open('gpurun_out/bench25_$d.json').read().strip().splitlines()[-1])
print('$d: fps=%.0f e2e=%.0f raw=%.0f'%(j['value'],j['e2e']['value'],j['e2e_raw']['value']), 'e2e gaps', j['e2e']['frame_ms'], 'raw gaps', j['e2e_raw']['frame_ms'])
"
done
