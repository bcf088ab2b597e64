mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest6.log
for cfg in "ldg 2" "tma 2" "ldg 1" "ldg 4" "tma 4"; do
  set -- $cfg
  B200_INTEGRATE_IMPL=$1 B200_RC_TPW=$2 timeout 600 python bench.py --steps 200 --cpu-steps 0 --harness-frames 0 --e2e-steps 8 > gpurun_out/bench6_$1_$2.json 2> gpurun_out/bench6_$1_$2.err
done
B200_INTEGRATE_IMPL=tma timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 600 --launch-count 10 -o gpurun_out/prof_frame_r1d python bench.py --steps 3 --warmup 1 --preroll 60 --e2e-steps 2 --cpu-steps 0 --harness-frames 0 > gpurun_out/ncu_frame6.log 2>&1
for f in gpurun_out/bench6_*.json; do echo $f; python -c "
import json,sys
j=json.loads(open('$f').read().strip().splitlines()[-1])
print('fps=%.0f ms=%.3f'%(j['value'],j['ms_per_step']), {k:round(v*1000) for k,v in j['stage_ms'].items()}, 'int_us=%.1f'%j['roofline']['mean_launch_us'])
"; done
