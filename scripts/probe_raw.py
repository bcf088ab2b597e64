"""GPU probe: per-stage timings of the raw-frame pipelined path vs the float-depth path (diagnostics only)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynslam_b200 import engine as E, synth

W, H = synth.KITTI_W, synth.KITTI_H
scene_s = synth.StreetScene(seed=6, length_m=200.0)
N0, N1 = 40, 12
frames = [synth.kitti_frame(scene_s, f) for f in range(N0 + 2 * N1)]
dev = torch.device("cuda:0")
scene = E.Scene(E.SceneParams(), 0x60000, 0x100000, 0x80000, device="cuda:0")
eng = E.Engine(scene, (W, H))
reco = E.SceneReconstructionEngine(eng)
rs = E.VisualisationEngine(eng, scene).CreateRenderState((W, H))
reco.ResetScene(scene)
points = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
normals = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
torch.cuda.synchronize()
for f in frames[:N0]:
    v = E.View(torch.from_numpy(f[0]).to(dev), torch.from_numpy(f[1]).to(dev), f[2], f[3])
    torch.cuda.synchronize()
    eng.process_frame_async(rs, v, points, normals, decay=(1, 200), raycast=True)
eng.sync(rs)
print("after preroll visible", rs.noVisibleBlocks)
ev = E.View(torch.zeros((H, W), dtype=torch.float32, device=dev), torch.zeros((H, W, 4), dtype=torch.uint8, device=dev), frames[0][2], frames[0][3])
out = [torch.zeros(H * W * 4, dtype=torch.uint8).pin_memory() for _ in range(2)]
calib = E.make_view_calib()
eng.set_timing(1)
torch.cuda.synchronize()
for mode in ("float", "raw"):
    base = N0 if mode == "float" else N0 + N1
    for i in range(N1):
        f = frames[base + i]
        hd = torch.from_numpy(f[0]).pin_memory()
        hr = torch.from_numpy(np.round(f[0] * 1000.0).astype(np.int16)).pin_memory()
        hc = torch.from_numpy(f[1]).pin_memory()
        ev.set_pose(f[2])
        t0 = time.perf_counter()
        if mode == "float":
            eng.host_frame_submit(rs, ev, hd, hc, points, normals, decay=(1, 200), h_out=out[0], slot=0)
        else:
            eng.host_frame_submit_raw(rs, ev, hr, hc, calib, points, normals, decay=(1, 200), h_out=out[0], slot=0)
        eng.host_frame_wait(0)
        eng.sync(rs)
        dt = (time.perf_counter() - t0) * 1e6
        st = eng.stats()
        print(mode, i, "wall_us %.0f" % dt, "alloc %.0f int %.0f exp %.0f ray %.0f decay %.0f total %.0f" % tuple(
            1000 * x for x in (st.ms_allocate, st.ms_integrate, st.ms_expected, st.ms_raycast, st.ms_decay, st.ms_total)),
            "visible", rs.noVisibleBlocks, "integrated", st.noIntegratedBlocks)
# UpdateView alone
vb = E.ViewBuilder(eng, calib)
raw = torch.from_numpy(np.round(frames[N0][0] * 1000.0).astype(np.int16)).to(dev)
dep = torch.zeros((H, W), dtype=torch.float32, device=dev)
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        vb.UpdateView(dep, raw, sync=False)
    eng.sync(rs)
    print("UpdateView device-only us/call %.1f" % ((time.perf_counter() - t0) / 50 * 1e6))
print("valid px", int((dep > 0).sum().item()), "min/max", float(dep[dep > 0].min()), float(dep.max()))
