"""GPU probe: launch trace (b200_set_timing(3)) of fused frames — where the frame's time goes, gaps included."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynslam_b200 import engine as E, synth

W, H = synth.KITTI_W, synth.KITTI_H
street = synth.StreetScene(seed=6, length_m=200.0)
N0, N1 = 60, 6
DECAY = (1, 30)
frames = [synth.kitti_frame(street, f) for f in range(N0 + N1)]
dev = torch.device("cuda:0")
scene = E.Scene(E.SceneParams(), 0x60000, 0x100000, 0x80000, device="cuda:0")
eng = E.Engine(scene, (W, H))
reco = E.SceneReconstructionEngine(eng)
rs = E.VisualisationEngine(eng, scene).CreateRenderState((W, H))
reco.ResetScene(scene)
points = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
normals = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
views = [E.View(torch.from_numpy(f[0]).to(dev), torch.from_numpy(f[1]).to(dev), f[2], f[3]) for f in frames]
flush = torch.zeros(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
for v in views[:N0]:
    eng.process_frame_async(rs, v, points, normals, decay=DECAY, raycast=True)
eng.sync(rs)
print("visible", rs.noVisibleBlocks)
for i, v in enumerate(views[N0:]):
    flush.add_(1); torch.cuda.synchronize()
    eng.set_timing(3)
    eng.process_frame_async(rs, v, points, normals, decay=DECAY, raycast=True)
    eng.sync(rs)
    tr = eng.trace()
    import ctypes as C
    dbg = (C.c_uint64 * (64 + 5120))()
    eng.lib.b200_diag_read_debug(eng.h, dbg, 64 + 5120)
    eng.set_timing(0)
    if i >= 2:
        print("--- frame", i)
        t0 = min(dbg[s * 8] for s in range(4) if dbg[s * 8])
        for s in range(4):
            print("  k_serve_list CTA slot", s, "phase stamps (us):", ["%.1f" % ((dbg[s * 8 + k] - t0) / 1000.0) if dbg[s * 8 + k] else "-" for k in range(8)])
        for s_ in range(4):
            print("  k_serve_list CTA slot", s_, "per-entry phase, first round (us): entry read, projected, group boxes, warp boxes, loop end:", ["%.1f" % ((dbg[32 + s_ * 8 + k] - t0) / 1000.0) if dbg[32 + s_ * 8 + k] else "-" for k in range(5)])
        cnt = [int(dbg[64 + 3 * 1024 + t]) for t in range(192)]
        pub = [(dbg[64 + 1024 + t] - t0) / 1000.0 for t in range(192)]
        srv = [(dbg[64 + 4 * 1024 + t] - t0) / 1000.0 if dbg[64 + 4 * 1024 + t] else -1 for t in range(192)]
        off = [(dbg[64 + 2 * 1024 + t] - t0) / 1000.0 for t in range(192)]
        print("  listed entries per tile: ordered part mean %.1f max %d; excess part:" % (np.mean(cnt[:128]), max(cnt[:128])), cnt[128:])
        print("  excess-part tiles, count published (us):", ["%.1f" % v for v in pub[128:]])
        print("  tiles by service end (us), slowest 12:", sorted([("%.1f" % v, t) for t, v in enumerate(srv)], key=lambda a: -float(a[0]))[:12])
        print("  ordered tiles, offset known (us), slowest 8:", sorted([("%.1f" % v, t) for t, v in enumerate(off[:128])], key=lambda a: -float(a[0]))[:8])
        for k, what in enumerate(["tile start", "list count published", "list offset known"]):
            v = np.array([dbg[64 + k * 1024 + t] for t in range(1024)], dtype=np.float64)
            v = (v[v > 0] - t0) / 1000.0
            if len(v):
                print("  per tile, %-22s n %4d  min %5.1f  p50 %5.1f  p90 %5.1f  max %5.1f (tile %d)" % (what, len(v), v.min(), np.percentile(v, 50), np.percentile(v, 90), v.max(), int(v.argmax())))
        for name, a, b in tr:
            print("%-22s start %7.1f  end %7.1f  dur %6.1f" % (name, a, b, b - a))

# UpdateView (five passes) under the trace
calib = E.make_view_calib()
vb = E.ViewBuilder(eng, calib)
raw = torch.from_numpy(np.round(frames[N0][0] * 1000.0).astype(np.int16)).to(dev)
dep = torch.zeros((H, W), dtype=torch.float32, device=dev)
vb.UpdateView(dep, raw)
for i in range(3):
    flush.add_(1); torch.cuda.synchronize()
    eng.set_timing(3)
    vb.UpdateView(dep, raw, sync=False)
    eng.sync(rs)
    tr = eng.trace()
    eng.set_timing(0)
print("--- UpdateView")
for name, a, b in tr:
    print("%-22s start %7.1f  end %7.1f  dur %6.1f" % (name, a, b, b - a))

import bench
print(bench.run_frames_ops(0))
