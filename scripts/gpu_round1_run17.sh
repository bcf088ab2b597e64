mkdir -p gpurun_out
timeout 600 python scripts/probe_trace.py 2>&1 | tail -90 | tee gpurun_out/probe17.log
