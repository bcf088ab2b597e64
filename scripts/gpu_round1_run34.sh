# ncu evidence: every kernel of one steady-state frame (cudaProfilerStart/Stop around timed step 1), full set + durations.
mkdir -p gpurun_out
COMMON="--preroll 215 --steps 3 --warmup 1 --e2e-steps 4 --e2e-raw-steps 0 --cpu-steps 0 --harness-frames 0 --hires-frames 0 --profile-step 1"
timeout 170 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_ \
   -o gpurun_out/prof_frame_r1h python bench.py $COMMON > gpurun_out/ncu_frame.log 2>&1
ls -la gpurun_out/*.ncu-rep
grep -c "==PROF== Profiling" gpurun_out/ncu_frame.log
tail -c 600 gpurun_out/ncu_frame.log
