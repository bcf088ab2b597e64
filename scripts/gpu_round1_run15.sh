mkdir -p gpurun_out
timeout 600 python scripts/probe_raw.py 2>&1 | tail -40 | tee gpurun_out/probe15.log
timeout 900 python -m pytest tests/test_gpu_view.py -m gpu -q -s 2>&1 | tail -8 | tee gpurun_out/pytest15.log
