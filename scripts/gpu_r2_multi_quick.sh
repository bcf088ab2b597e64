#!/bin/bash
# quick N-GPU check of the exchange transports: push (default) vs nccl, timed loop only
set -u
N=${1:-2}
mkdir -p gpurun_out
COMMON="--gpus $N --steps 200 --warmup 5 --harness-frames 0 --hires-frames 0 --decay-blocks 0 --no-parity-check --cpu-steps 0 --e2e-steps 60"
for v in push; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py $COMMON \
     > gpurun_out/xch_n${N}_$v.json 2> gpurun_out/xch_n${N}_$v.err
  echo "rc=$?"; grep -i "error\|composite\|own timed" gpurun_out/xch_n${N}_$v.err | head -8
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/xch_n${N}_$v.json").read().strip().splitlines()[-1])
    print("$v", "value %.0f" % d["value"], "ms_per_step %.4f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"])
except Exception as ex:
    print("$v", "unreadable:", ex)
PY
done
