"""CPU statistics of the raycast march on the KITTI-shaped stream (host build of raycast_ray.cuh): steps through missing blocks
and interpolated samples per ray, and per 8x4-pixel warp (max over lanes = what the warp executes). Design aid only."""
import ctypes as C, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dynslam_b200 import synth
from tests import hostlib as H

HERE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "hostcheck")
subprocess.run(["make", "-C", HERE, "all"], check=True, stdout=subprocess.DEVNULL)
lib = C.CDLL(os.path.join(HERE, "libhostcheck.so"))
L = H.oracle()
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40
w, h = synth.KITTI_W, synth.KITTI_H
p = H.SceneParams()
vol = H.HostVolume(0x60000, 0x100000, 0x80000, w, h, p)
street = synth.StreetScene(seed=6, length_m=200.0)
for f in range(frames):
    depth, rgb, M, proj = synth.kitti_frame(street, f)
    hv = H.make_view(depth, rgb, M, proj)
    assert L.oracle_allocate_from_depth(vol.engine, C.byref(vol.scene), C.byref(vol.rs), C.byref(hv), 0, 0) == 0
    L.oracle_integrate(vol.engine, C.byref(vol.scene), C.byref(vol.rs), C.byref(hv), 0)
cam = H.make_camera(M, proj)
L.oracle_expected_depths(C.byref(vol.scene), C.byref(vol.rs), C.byref(cam))
ne = np.zeros((h, w), np.int32); nf = np.zeros((h, w), np.int32)
vp = C.c_void_p
lib.hostcheck_raycast_stats.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_float, vp, vp, vp]
lib.hostcheck_raycast_stats(H.vptr(vol.voxels), H.vptr(vol.hash), 0x100000, w, h, hv.invM_d, hv.proj_d, p.voxelSize, p.mu, H.vptr(vol.minmax), H.vptr(ne), H.vptr(nf))
print("rays", w * h, "with any step", int(((ne + nf) > 0).sum()))
print("per ray: empty steps mean %.2f max %d; interpolated samples mean %.2f max %d" % (ne.mean(), ne.max(), nf.mean(), nf.max()))
# per warp (8x4 tile)
H4, W8 = (h + 3) // 4, (w + 7) // 8
pad = lambda a: np.pad(a, ((0, H4 * 4 - h), (0, W8 * 8 - w)))
t = lambda a: pad(a).reshape(H4, 4, W8, 8).transpose(0, 2, 1, 3).reshape(H4, W8, 32)
te, tf, tt = t(ne), t(nf), t(ne + nf)
print("warps", H4 * W8, "active", int((tt.max(-1) > 0).sum()))
print("per warp: max total steps mean %.1f p50 %.0f p90 %.0f p99 %.0f max %d" % ((tt.max(-1).mean(),) + tuple(np.percentile(tt.max(-1), [50, 90, 99])) + (tt.max(-1).max(),)))
print("  sum over warps of max(total) = %d ; sum over rays of total / 32 = %.0f  (lane utilisation of the march %.2f)" % (tt.max(-1).sum(), tt.sum() / 32.0, tt.sum() / 32.0 / tt.max(-1).sum()))
print("  sum over warps of max(empty) = %d, of max(found) = %d" % (te.max(-1).sum(), tf.max(-1).sum()))
rows = tt.max(-1).sum(1)
print("  heaviest tile rows:", np.argsort(-rows)[:8], rows[np.argsort(-rows)[:8]])
hist = np.bincount(np.minimum(ne.ravel(), 200) // 10)
print("  empty steps per ray histogram (bins of 10, last = 200+):", hist.tolist())

# chain walks of the neighbourhood-cache march (cast_ray_nbr), per interpolated sample
out = np.zeros((h, w, 4), np.float32)
lib.hostcheck_raycast.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_float, vp, vp, C.c_int]
lib.hostcheck_raycast_walks.restype = C.c_long
lib.hostcheck_raycast_walks(1)
lib.hostcheck_raycast(H.vptr(vol.voxels), H.vptr(vol.hash), 0x100000, w, h, hv.invM_d, hv.proj_d, p.voxelSize, p.mu, H.vptr(vol.minmax), H.vptr(out), 1)
walks = lib.hostcheck_raycast_walks(1)
print("cast_ray_nbr: %d chain walks = %.2f per step, %.2f per interpolated sample (per ray, not per warp)" % (walks, walks / max((ne + nf).sum(), 1), walks / max(nf.sum(), 1)))
