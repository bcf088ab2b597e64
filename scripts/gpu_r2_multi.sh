#!/bin/bash
# N-GPU run of the multi-volume bench (configs[2]) next to a 1-GPU run on the same box. usage: gpu_r2_multi.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 200 --warmup 5 \
   > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; tail -3 gpurun_out/bench_n$N.err
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 5 --harness-frames 0 --hires-frames 0 --decay-blocks 0 --no-parity-check --cpu-steps 0 \
   > gpurun_out/bench_n1_same_box.json 2> gpurun_out/bench_n1_same_box.err
echo "rc=$?"; tail -2 gpurun_out/bench_n1_same_box.err
python - <<PY
import json
for f in ["gpurun_out/bench_n$N.json", "gpurun_out/bench_n1_same_box.json"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ["value", "n_gpus", "ms_per_step", "scaling"]}, json.dumps(d.get("exchange"))[:600], json.dumps(d.get("e2e"))[:300])
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
