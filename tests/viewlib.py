"""Shared inputs for the view-builder tests: raw int16 depth frames shaped like DynSLAM's input
(millimetres, 0 = invalid; DS/InfiniTamDriver.cpp:52,77) plus the edge values convertDepthAffineToFloat
and filterDepth distinguish (negative, > 32000, isolated valid pixels, holes next to the image border)."""
import numpy as np

from dynslam_b200 import synth


def raw_kitti_frame(frame=3, scale=1.0, seed=0):
    scene = synth.StreetScene(seed=6)
    depth, rgb, _, _ = synth.kitti_frame(scene, frame, scale=scale)
    raw = np.round(depth * 1000.0).astype(np.int16)          # metres -> int mm, as DynSLAM feeds InfiniTAM
    rng = np.random.default_rng(seed)
    h, w = raw.shape
    for _ in range(40):                                       # holes, some touching the border
        y, x = rng.integers(0, h), rng.integers(0, w)
        raw[max(0, y - 3):y + 4, max(0, x - 5):x + 6] = 0
    raw[rng.integers(0, h, 30), rng.integers(0, w, 30)] = -5            # negative -> invalid
    raw[rng.integers(0, h, 30), rng.integers(0, w, 30)] = 32001         # > 32000 -> invalid
    raw[rng.integers(0, h, 30), rng.integers(0, w, 30)] = 32000         # largest valid
    raw[0:2, 10:40] = 1500                                              # valid values on the 2-pixel border
    raw[:, 0] = 0
    return raw, rgb


def raw_noise_frame(w=97, h=61, seed=1):
    """Small ragged-size frame (not a multiple of any tile) with a noisy plane and 10 % holes."""
    rng = np.random.default_rng(seed)
    z = 2000 + 300 * np.sin(np.arange(w)[None, :] * 0.21) + 200 * np.cos(np.arange(h)[:, None] * 0.17)
    raw = (z + rng.normal(0, 15, (h, w))).astype(np.int16)
    raw[rng.random((h, w)) < 0.10] = 0
    return raw
