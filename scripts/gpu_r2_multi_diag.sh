#!/bin/bash
# attribution of the per-step cost of configs[2] (N GPUs) over configs[1]: the same run with parts switched off
set -u
N=${1:-2}
mkdir -p gpurun_out
COMMON="--gpus $N --steps 200 --warmup 5 --harness-frames 0 --hires-frames 0 --decay-blocks 0 --no-parity-check --cpu-steps 0 --e2e-steps 8"
for v in full noxch noxch,norender noxch,norender,nosplit; do
  B200_BENCH_DIAG=$([ $v = full ] && echo none || echo $v) timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py $COMMON \
     > gpurun_out/diag_n${N}_$v.json 2> gpurun_out/diag_n${N}_$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/diag_n${N}_$v.json").read().strip().splitlines()[-1])
    print("$v", "value %.0f" % d["value"], "ms_per_step %.4f" % d["ms_per_step"], "stage", json.dumps(d["stage_ms"]))
except Exception as ex:
    print("$v", "unreadable:", ex)
PY
done
