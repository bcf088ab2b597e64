# ncu evidence for the V3 integrate kernel: one steady-state KITTI launch and one launch on the 4 mm stream.
mkdir -p gpurun_out
COMMON="--preroll 215 --steps 3 --warmup 1 --e2e-steps 4 --e2e-raw-steps 0 --cpu-steps 0 --harness-frames 0 --hires-frames 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_integrate --launch-skip 217 --launch-count 1 \
   -o gpurun_out/prof_integrate_v3_kitti python bench.py $COMMON > gpurun_out/ncu_v3_kitti.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_integrate --launch-skip 13 --launch-count 1 \
   -o gpurun_out/prof_integrate_v3_hires python scripts/profile_hires.py 16 > gpurun_out/ncu_v3_hires.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/ncu_v3_kitti.log gpurun_out/ncu_v3_hires.log
