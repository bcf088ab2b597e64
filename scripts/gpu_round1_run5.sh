mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest5.log
timeout 900 python bench.py --cpu-steps 0 --harness-frames 0 > gpurun_out/bench5_ldg.json 2> gpurun_out/bench5_ldg.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 600 --launch-count 12 -o gpurun_out/prof_frame_r1c python bench.py --steps 3 --warmup 1 --preroll 60 --e2e-steps 2 --cpu-steps 0 --harness-frames 0 > gpurun_out/ncu_frame5.log 2>&1
tail -2 gpurun_out/bench5_ldg.err; cat gpurun_out/bench5_ldg.json | head -c 3000
