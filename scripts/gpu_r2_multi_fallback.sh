#!/bin/bash
# 2-GPU check of the exchange set-up: the push transport, and the collective fall-back to the nccl transport when a rank's IPC
# open fails (simulated with B200_COMM_DIAG=8)
set -u
N=${1:-2}
mkdir -p gpurun_out
COMMON="--gpus $N --steps 100 --warmup 3 --preroll 60 --harness-frames 0 --hires-frames 0 --decay-blocks 0 --no-parity-check --cpu-steps 0 --e2e-steps 8"
for v in ${2:-0 8}; do
  B200_COMM_DIAG=$v timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py $COMMON \
     > gpurun_out/fb_n${N}_$v.json 2> gpurun_out/fb_n${N}_$v.err
  echo "diag=$v rc=$?"; grep -i "error\|composite:" gpurun_out/fb_n${N}_$v.err | head -4
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/fb_n${N}_$v.json").read().strip().splitlines()[-1])
    print("diag $v", "value %.0f" % d["value"], "ms_per_step %.4f" % d["ms_per_step"])
except Exception as ex:
    print("diag $v unreadable:", ex)
PY
done
