"""Shared inputs for the instance-splitting / compositing tests: a KITTI-shaped frame with
moving cars, the detections' bounding boxes and box-sized masks (copy mask = silhouette dilated, delete mask =
silhouette dilated further, as DynSLAM's segmentation provider produces them), including a box that sticks out
of the frame and overlapping detections."""
import ctypes as C

import numpy as np

from dynslam_b200 import abi, synth


def _dilate(m, r):
    out = m.copy()
    for dy in range(-r, r + 1):
        for dx in range(-r, r + 1):
            sh = np.zeros_like(m)
            ys, yd = (slice(dy, None), slice(0, m.shape[0] - dy)) if dy >= 0 else (slice(0, dy), slice(-dy, None))
            xs, xd = (slice(dx, None), slice(0, m.shape[1] - dx)) if dx >= 0 else (slice(0, dx), slice(-dx, None))
            sh[yd, xd] = m[ys, xs]
            out |= sh
    return out


def box_mask(full, pad=0):
    """full: bool [h, w] -> ((x0, y0, x1, y1) inclusive, uint8 box-sized mask). pad grows the box beyond the silhouette
    (and possibly beyond the frame: the mask rows/cols that fall outside are 0/1 noise the reference never reads)."""
    ys, xs = np.nonzero(full)
    h, w = full.shape
    x0, x1, y0, y1 = xs.min() - pad, xs.max() + pad, ys.min() - pad, ys.max() + pad
    data = np.zeros((y1 - y0 + 1, x1 - x0 + 1), np.uint8)
    cx0, cy0, cx1, cy1 = max(x0, 0), max(y0, 0), min(x1, w - 1), min(y1, h - 1)
    data[cy0 - y0:cy1 - y0 + 1, cx0 - x0:cx1 - x0 + 1] = full[cy0:cy1 + 1, cx0:cx1 + 1]
    data[0, :] |= 1 if (y0 < 0) else 0          # out-of-frame rows carry 1s: must be ignored (row/col range check)
    return (int(x0), int(y0), int(x1), int(y1)), data


def scene_with_cars(scale=0.5, frame=2, ncars=3):
    street = synth.StreetScene(seed=3, length_m=80.0)
    cars = [synth.MovingCar(i, seed=3) for i in range(ncars)]
    proj = synth.kitti_intrinsics() * np.float32(scale)
    w, h = int(round(synth.KITTI_W * scale)), int(round(synth.KITTI_H * scale))
    M = synth.kitti_pose(frame)
    depth, rgb, ident = street.render(M, w, h, float(proj[0]), float(proj[1]), float(proj[2]), float(proj[3]),
                                      extra_boxes=[c.box(frame) for c in cars], want_ids=True)
    dets = []
    for i in range(ncars):
        sil = ident == (1000 + i)
        if sil.sum() < 20:
            continue
        copy_box, copy_data = box_mask(_dilate(sil, 1), pad=2 if i == 0 else 0)
        del_box, del_data = box_mask(_dilate(sil, 3), pad=40 if i == 1 else 1)     # i == 1: box sticks out of the frame
        dets.append(dict(copy=(copy_box, copy_data), delete=(del_box, del_data)))
    return depth, rgb, dets


def host_mask(box, data):
    m = abi.Mask()
    m.x0, m.y0, m.x1, m.y1 = box
    m.d_data = data.ctypes.data
    m._keep = data
    return m


def host_ops(dets, actions, w, h):
    """b200_silhouette_op array with HOST pointers (for the oracle) + the destination arrays."""
    ops = (abi.SilhouetteOp * len(dets))()
    dests = []
    for k, (d, a) in enumerate(zip(dets, actions)):
        ops[k].action = a
        ops[k].copy_mask = host_mask(*d["copy"])
        ops[k].delete_mask = host_mask(*d["delete"])
        drgb = np.full((h, w, 4), 7, np.uint8)
        ddep = np.full((h, w), 7.0, np.float32)
        dests.append((drgb, ddep))
        if a == 2:
            ops[k].d_dest_rgb, ops[k].d_dest_depth = drgb.ctypes.data, ddep.ctypes.data
    ops._keep = dests
    return ops, dests
