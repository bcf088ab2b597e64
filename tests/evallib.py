"""Shared by the evaluation-consumer tests: a synthetic KITTI-like rig (velodyne -> camera transform, left / right colour
projections), LIDAR returns sampled from a depth image, and the ctypes plumbing for the oracle and the reference pin."""
import ctypes as C
import os

import numpy as np

from dynslam_b200 import abi
from tests import hostlib as H

EVALREF_SO = os.path.join(H.ROOT, "oracle", "_ref", "libevalref.so")


def rig(w, h, fx=721.5377, baseline=0.5371657):
    """KITTI-odometry-shaped calibration: Tr_velo_to_cam (x forward, y left, z up -> camera z forward, x right, y down), P2, P3."""
    cx, cy = w / 2.0 + 3.2, h / 2.0 - 5.1
    velo_to_cam = np.array([[7.5337e-03, -9.999714e-01, -6.16602e-04, -4.069766e-03],
                            [1.480249e-02, 7.280733e-04, -9.998902e-01, -7.631618e-02],
                            [9.998621e-01, 7.523790e-03, 1.480755e-02, -2.717806e-01],
                            [0.0, 0.0, 0.0, 1.0]], dtype=np.float64)
    pl = np.array([[fx, 0, cx, 4.485728e+01], [0, fx, cy, 2.163791e-01], [0, 0, 1, 2.745884e-03]], dtype=np.float64)
    pr = np.array([[fx, 0, cx, 4.485728e+01 - fx * baseline], [0, fx, cy, 2.163791e-01], [0, 0, 1, 2.745884e-03]], dtype=np.float64)
    return velo_to_cam, pl, pr, baseline


def params(w, h, min_depth=0.5, max_depth=30.0):
    v, pl, pr, b = rig(w, h)
    p = abi.EvalParams()
    p.velo_to_cam[:] = list(v.T.reshape(-1))
    p.proj_left[:] = list(pl.T.reshape(-1))
    p.proj_right[:] = list(pr.T.reshape(-1))
    p.baseline_m = b
    p.left_focal_length_px = float(np.float32(pl[0, 0]))
    p.min_depth_m, p.max_depth_m = min_depth, max_depth
    p.frame_width, p.frame_height = w, h
    return p, (v, pl, pr, b)


def lidar_cloud(n, seed, rigt, w, h):
    """n returns in velodyne coordinates: most inside the camera frustum, some behind / outside / too far, a few exactly on
    rounding boundaries of the pixel grid"""
    rng = np.random.default_rng(seed)
    v, pl, pr, _ = rigt
    z = rng.uniform(0.2, 45.0, n)
    u = rng.uniform(-40, w + 40, n)
    r = rng.uniform(-30, h + 30, n)
    x = (u - pl[0, 2]) * z / pl[0, 0]
    y = (r - pl[1, 2]) * z / pl[1, 1]
    cam = np.stack([x, y, z, np.ones(n)], 1)
    cam[::97, 2] *= -1.0                                    # behind the camera
    velo = (np.linalg.inv(v) @ cam.T).T
    pts = np.zeros((n, 4), np.float32)
    pts[:, :3] = velo[:, :3].astype(np.float32)
    pts[:, 3] = rng.uniform(0, 1, n).astype(np.float32)     # reflectance (ignored)
    return pts


def depth_images(w, h, seed):
    """rendered depth (float metres, zeros = missing) and input depth (int16 mm, zeros = missing), correlated, with holes"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = 4.0 + 20.0 * (1.0 - yy / h) + 2.0 * np.sin(xx / 37.0)
    rendered = (base + rng.normal(0, 0.15, (h, w))).astype(np.float32)
    inp = np.clip(np.round((base + rng.normal(0, 0.4, (h, w))) * 1000.0), 0, 32000).astype(np.int16)
    rendered[rng.uniform(size=(h, w)) < 0.15] = 0.0
    inp[rng.uniform(size=(h, w)) < 0.10] = 0
    rendered[rng.uniform(size=(h, w)) < 0.01] = 5e-6       # |depth| < 1e-5 counts as missing
    return rendered, inp


def callbacks_array(cbs):
    return (abi.EvalCallback * len(cbs))(*[abi.EvalCallback(float(d), int(bool(c)), int(bool(k))) for d, c, k in cbs])


def run_oracle(p, pts, rendered, inp, cbs, association=None, with_dynamic=False):
    L = H.oracle()
    n = len(cbs)
    out_s, out_d, summ = (abi.EvalResult * n)(), ((abi.EvalResult * n)() if with_dynamic else None), abi.EvalSummary()
    rc = L.oracle_evaluate_depth(C.byref(p), H.vptr(pts), len(pts), H.vptr(rendered), H.vptr(inp), H.vptr(association) if association is not None else None,
                                 callbacks_array(cbs), n, out_s, out_d, C.byref(summ))
    return rc, [r.as_dict() for r in out_s], ([r.as_dict() for r in out_d] if with_dynamic else None), summ


def evalref_available():
    return os.path.exists(EVALREF_SO)


def run_reference(p, pts, rendered, inp, cbs, association=None, with_dynamic=False):
    L = C.CDLL(EVALREF_SO)
    P, vp = C.POINTER, C.c_void_p
    L.ref_evaluate_depth.argtypes = [P(abi.EvalParams), vp, C.c_int, vp, vp, vp, P(abi.EvalCallback), C.c_int, P(abi.EvalResult), P(abi.EvalResult), P(C.c_long)]
    n = len(cbs)
    out_s, out_d, sk = (abi.EvalResult * n)(), ((abi.EvalResult * n)() if with_dynamic else None), C.c_long(0)
    rc = L.ref_evaluate_depth(C.byref(p), H.vptr(pts), len(pts), H.vptr(rendered), H.vptr(inp), H.vptr(association) if association is not None else None,
                              callbacks_array(cbs), n, out_s, out_d, C.byref(sk))
    return rc, [r.as_dict() for r in out_s], ([r.as_dict() for r in out_d] if with_dynamic else None), sk.value
