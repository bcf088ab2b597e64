"""Profiling target: builds the 4 mm roofline-stress map (bench.run_hires) so that ncu can capture one steady-state
IntegrateIntoScene launch on it (no timing here)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

print(bench.run_hires(0, int(sys.argv[1]) if len(sys.argv) > 1 else 16))
