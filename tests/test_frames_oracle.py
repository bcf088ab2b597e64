"""CPU: known-answer checks of oracle/frames_oracle.c (instance splitting + compositing). The reference functions
(DS/InstRecLib/InstanceReconstructor.cpp:59-170, :850-905) cannot be compiled here (OpenCV/Eigen/Pangolin), so the
oracle is PARITY UNPINNED; these hand-computed cases fix the statement-level semantics it restates."""
import ctypes as C

import numpy as np

from dynslam_b200 import abi
from tests import frameslib as F
from tests import hostlib as H


def test_process_and_remove_silhouette_known_answers():
    L = H.oracle()
    w, h = 8, 6
    rgb = np.arange(w * h * 4, dtype=np.uint8).reshape(h, w, 4)
    depth = (np.arange(w * h, dtype=np.float32).reshape(h, w) + 1.0)
    rgb0, depth0 = rgb.copy(), depth.copy()
    copy = ((2, 1, 4, 3), np.array([[1, 1, 0], [1, 1, 1], [0, 1, 2]], np.uint8))          # 2 is not "inside" (== 1 test)
    dele = ((1, 0, 5, 4), np.zeros((5, 5), np.uint8))
    dele[1][1:4, 1:4] = 1                                                                  # frame x 2..4, y 1..3
    ops, dests = F.host_ops([dict(copy=copy, delete=dele)], [2], w, h)
    L.oracle_process_silhouettes(H.vptr(rgb), H.vptr(depth), w, h, ops, 1)
    drgb, ddep = dests[0]
    inside = np.zeros((h, w), bool)
    inside[1, 2] = inside[1, 3] = inside[2, 2] = inside[2, 3] = inside[2, 4] = inside[3, 3] = True
    assert np.array_equal(drgb[inside], rgb0[inside]) and np.array_equal(ddep[inside], depth0[inside])
    assert (drgb[~inside] == 255).all() and (ddep[~inside] == 0).all()                    # memset 255 / 0 everywhere else
    cut = np.zeros((h, w), bool); cut[1:4, 2:5] = True
    assert (rgb[cut] == 0).all() and (depth[cut] == 0).all()
    assert np.array_equal(rgb[~cut], rgb0[~cut]) and np.array_equal(depth[~cut], depth0[~cut])
    # a second detection copying from the already blanked area gets zeros (sequential semantics)
    rgb, depth = rgb0.copy(), depth0.copy()
    ops, dests = F.host_ops([dict(copy=copy, delete=dele), dict(copy=copy, delete=dele)], [1, 2], w, h)
    L.oracle_process_silhouettes(H.vptr(rgb), H.vptr(depth), w, h, ops, 2)
    assert (dests[1][1][inside] == 0).all() and (dests[1][0][inside] == 0).all()
    # action 0 touches nothing
    rgb, depth = rgb0.copy(), depth0.copy()
    ops, _ = F.host_ops([dict(copy=copy, delete=dele)], [0], w, h)
    L.oracle_process_silhouettes(H.vptr(rgb), H.vptr(depth), w, h, ops, 1)
    assert np.array_equal(rgb, rgb0) and np.array_equal(depth, depth0)


def test_composite_known_answers():
    L = H.oracle()
    t = np.array([0.0, 2.0, 3.0, 0.0, 5.0], np.float32)
    s = np.array([1.0, 0.0, 2.5, 0.0, 6.0], np.float32)
    L.oracle_composite_depth(H.vptr(t), H.vptr(s), 5)
    assert t.tolist() == [1.0, 2.0, 2.5, 0.0, 5.0]
    tc = np.array([[10, 20, 30, 40]] * 4, np.uint8)
    td = np.array([0.0, 4.0, 4.0, 4.0], np.float32)
    sc = np.array([[100, 200, 250, 9]] * 4, np.uint8)
    sd = np.array([3.0, 3.0, 5.0, 0.0], np.float32)
    tint = (C.c_int32 * 4)(0x1f, 0x77, 0xb4, 255)
    L.oracle_composite_color(H.vptr(tc), H.vptr(td), H.vptr(sc), H.vptr(sd), 4, tint, 1.0)
    # tint_strength 1: col_strength = 0.5 -> (100*.5+31, 200*.5+119, min(255, 250*.5+180)) ; alpha untouched
    assert tc[0].tolist() == [81, 219, 255, 40] and tc[1].tolist() == [81, 219, 255, 40]
    assert tc[2].tolist() == [10, 20, 30, 40] and tc[3].tolist() == [10, 20, 30, 40]
    assert td.tolist() == [3.0, 3.0, 4.0, 4.0]
    # dimming: uchar(c * (1.0 - 0.10f)) in double: 10*0.9 = 8.99999997 -> 8 ; 200 -> 179 (0.10f > 0.1)
    oc = np.array([[10, 200, 255, 77]], np.uint8)
    od = np.array([1.0], np.float32)
    layers = (abi.InstanceLayer * 1)()
    L.oracle_composite_instances(H.vptr(oc), H.vptr(od), 1, layers, 0, 0.10, 1.0)
    assert oc[0].tolist() == [8, 179, 229, 77]
