mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus2.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 200 --warmup 5 --cpu-steps 0 > gpurun_out/bench11_2gpu.json 2> gpurun_out/bench11_2gpu.err
tail -5 gpurun_out/bench11_2gpu.err
python -c "
import json
j=json.loads(open('gpurun_out/bench11_2gpu.json').read().strip().splitlines()[-1])
print('n=%d fps=%.0f ms=%.3f e2e=%.0f'%(j['n_gpus'],j['value'],j['ms_per_step'],j['e2e']['value']), {k:round(v*1000) for k,v in j['stage_ms'].items()})
"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 8 > gpurun_out/bench11_2gpu_ref.json 2> gpurun_out/bench11_2gpu_ref.err
cut -c1-200 gpurun_out/bench11_2gpu_ref.json
