"""GPU probe: launch trace of the PIPELINED host-frame path (copies included)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynslam_b200 import engine as E, synth

W, H = synth.KITTI_W, synth.KITTI_H
street = synth.StreetScene(seed=6, length_m=200.0)
N0, N1 = 40, 24
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
views = [E.View(torch.from_numpy(f[0]).to(dev), torch.from_numpy(f[1]).to(dev), f[2], f[3]) for f in frames[:N0]]
torch.cuda.synchronize()
for v in views:
    eng.process_frame_async(rs, v, points, normals, decay=DECAY, raycast=True)
eng.sync(rs)
hd = [torch.from_numpy(f[0]).pin_memory() for f in frames[N0:]]
hc = [torch.from_numpy(f[1]).pin_memory() for f in frames[N0:]]
out = [torch.zeros(H * W * 4, dtype=torch.uint8).pin_memory() for _ in range(2)]
ev = E.View(torch.zeros((H, W), dtype=torch.float32, device=dev), torch.zeros((H, W, 4), dtype=torch.uint8, device=dev), frames[0][2], frames[0][3])
for mode in (0, 3):
    eng.set_timing(mode)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(N1):
        slot = i & 1
        eng.host_frame_wait(slot)
        ev.set_pose(frames[N0 + i][2])
        eng.host_frame_submit(rs, ev, hd[i], hc[i], points, normals, decay=DECAY, h_out=out[slot], slot=slot)
    eng.host_frame_wait(0); eng.host_frame_wait(1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("mode", mode, "pipelined us/frame %.1f" % (dt / N1 * 1e6))
    if mode == 3:
        tr = eng.trace()
        t_first = None
        for name, a, b in tr[-42:]:
            print("%-22s start %8.1f  end %8.1f  dur %7.1f" % (name, a, b, b - a))
    eng.set_timing(0)
