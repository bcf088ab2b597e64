mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest2.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_ldg.json 2> gpurun_out/bench_ldg.err
B200_INTEGRATE_IMPL=tma timeout 600 python bench.py --steps 100 --warmup 5 --cpu-steps 0 > gpurun_out/bench_tma.json 2> gpurun_out/bench_tma.err
timeout 600 python bench.py --steps 100 --warmup 5 --cpu-steps 0 --no-flush-l2 > gpurun_out/bench_ldg_warm.json 2> gpurun_out/bench_ldg_warm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 1 --preroll 40 --e2e-steps 2 --cpu-steps 0 > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 40 -c 3 -o gpurun_out/prof_integrate_ldg python bench.py --steps 3 --warmup 1 --preroll 40 --e2e-steps 2 --cpu-steps 0 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/bench_ldg.err; cat gpurun_out/bench_ldg.json | head -c 3000
