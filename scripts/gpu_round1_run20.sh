mkdir -p gpurun_out
for r in 48 56; do echo "=== regs $r"; B200_INTEGRATE_REGS=$r timeout 600 python scripts/probe_e2e.py 2>&1 | tail -46; done | tee gpurun_out/probe20.log
