mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest4.log
timeout 900 python bench.py --cpu-steps 0 --harness-frames 0 > gpurun_out/bench4_ldg.json 2> gpurun_out/bench4_ldg.err
B200_INTEGRATE_IMPL=tma timeout 600 python bench.py --cpu-steps 0 --harness-frames 0 > gpurun_out/bench4_tma.json 2> gpurun_out/bench4_tma.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 540 --launch-count 12 -o gpurun_out/prof_frame_r1b python bench.py --steps 3 --warmup 1 --preroll 60 --e2e-steps 2 --cpu-steps 0 --harness-frames 0 > gpurun_out/ncu_frame4.log 2>&1
tail -2 gpurun_out/bench4_ldg.err; cat gpurun_out/bench4_ldg.json | head -c 3000
