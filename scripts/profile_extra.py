"""Runs the view-builder and frame-op kernels a few times (profiling target for ncu; no timing here)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynslam_b200 import engine as E, synth
import bench

W, H = synth.KITTI_W, synth.KITTI_H
street = synth.StreetScene(seed=6, length_m=60.0)
depth, rgb, M, proj = synth.kitti_frame(street, 3)
dev = torch.device("cuda:0")
eng = E.Engine(E.Scene(E.SceneParams(), 2048, 0x800, 0x400, device="cuda:0"), (W, H))
vb = E.ViewBuilder(eng, E.make_view_calib(intrinsics_d=tuple(proj), modelSensorNoise=True))
raw = torch.from_numpy(np.round(depth * 1000.0).astype(np.int16)).to(dev)
out = torch.zeros((H, W), dtype=torch.float32, device=dev)
nrm = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
sig = torch.zeros((H, W), dtype=torch.float32, device=dev)
for _ in range(3):
    vb.UpdateView(out, raw, nrm, sig)
print(bench.run_frames_ops(0, iters=3))
