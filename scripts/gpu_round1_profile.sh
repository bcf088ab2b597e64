# final ncu evidence of the round: launch list of the bench command + full captures of one steady-state frame and of the
# view-builder / frame-op kernels. Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
COMMON="--preroll 215 --steps 3 --warmup 1 --e2e-steps 4 --e2e-raw-steps 0 --cpu-steps 0 --harness-frames 0 --hires-frames 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --launch-skip 2190 --launch-count 84 --csv \
   --log-file gpurun_out/r01g_launches.csv python bench.py $COMMON > gpurun_out/ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 2202 --launch-count 13 \
   -o gpurun_out/prof_frame_r1g python bench.py $COMMON > gpurun_out/ncu_frame.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_filter_pass|k_normal_weight|k_process_silhouettes|k_composite_layers" \
   --launch-skip 12 --launch-count 8 -o gpurun_out/prof_extra_r1g python scripts/profile_extra.py > gpurun_out/ncu_extra.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r01g_launches.csv
tail -3 gpurun_out/ncu_launches.log gpurun_out/ncu_frame.log gpurun_out/ncu_extra.log
