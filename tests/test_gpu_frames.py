"""GPU: instance frame splitting and render compositing (dynslam_b200/csrc/frames.cu)
against oracle/frames_oracle.c, bit for bit (byte / integer / compare-select work)."""
import ctypes as C

import numpy as np
import pytest
import torch

from dynslam_b200 import abi, engine as E
from tests import frameslib as F
from tests import hostlib as H
from tests import parity as P

pytestmark = pytest.mark.gpu


def _engine():
    return E.Engine(E.Scene(E.SceneParams(), 2048, 0x800, 0x400, "cuda:0"), (64, 48))


@pytest.mark.parametrize("actions", [(2, 2, 2), (1, 2, 0), (2, 1, 2), (0, 0, 0)])
def test_process_silhouettes_equals_oracle(actions):
    depth, rgb, dets = F.scene_with_cars()
    assert len(dets) == 3
    h, w = depth.shape
    L = H.oracle()
    # oracle (host pointers)
    o_rgb, o_depth = rgb.copy(), depth.copy()
    ops, dests = F.host_ops(dets, actions, w, h)
    L.oracle_process_silhouettes(H.vptr(o_rgb), H.vptr(o_depth), w, h, ops, len(dets))
    # CUDA
    eng = _engine()
    fr = E.InstanceFrames(eng)
    d_rgb, d_depth = torch.from_numpy(rgb).cuda(), torch.from_numpy(depth).cuda()
    d_ops, d_dests = [], []
    for det, a in zip(dets, actions):
        cm = E.make_mask(det["copy"][0], torch.from_numpy(det["copy"][1]).cuda())
        dm = E.make_mask(det["delete"][0], torch.from_numpy(det["delete"][1]).cuda())
        drgb = torch.full((h, w, 4), 7, dtype=torch.uint8, device="cuda")
        ddep = torch.full((h, w), 7.0, dtype=torch.float32, device="cuda")
        d_dests.append((drgb, ddep))
        d_ops.append((a, cm, dm, drgb if a == 2 else None, ddep if a == 2 else None))
    fr.ProcessSilhouettes(d_rgb, d_depth, d_ops)
    P._cmp("main rgb", d_rgb.cpu().numpy(), o_rgb)
    P._cmp("main depth", d_depth.cpu().numpy(), o_depth)
    for k, a in enumerate(actions):
        P._cmp(f"instance {k} rgb", d_dests[k][0].cpu().numpy(), dests[k][0])
        P._cmp(f"instance {k} depth", d_dests[k][1].cpu().numpy(), dests[k][1])
    if 2 in actions:
        k = actions.index(2)
        assert (dests[k][1] > 0).sum() > 50          # the silhouette really was copied
    if any(actions):
        assert (o_depth != depth).any()


def test_process_silhouettes_more_ops_than_one_launch_holds():
    """> 24 ops are applied in chunks; order across the chunk boundary must still be the list order."""
    depth, rgb, dets = F.scene_with_cars()
    h, w = depth.shape
    many = [dets[i % len(dets)] for i in range(30)]
    actions = [2 if i in (0, 29) else 1 for i in range(30)]
    L = H.oracle()
    o_rgb, o_depth = rgb.copy(), depth.copy()
    ops, dests = F.host_ops(many, actions, w, h)
    L.oracle_process_silhouettes(H.vptr(o_rgb), H.vptr(o_depth), w, h, ops, 30)
    eng = _engine()
    fr = E.InstanceFrames(eng)
    d_rgb, d_depth = torch.from_numpy(rgb).cuda(), torch.from_numpy(depth).cuda()
    d_ops, d_dests = [], {}
    for i, (det, a) in enumerate(zip(many, actions)):
        cm = E.make_mask(det["copy"][0], torch.from_numpy(det["copy"][1]).cuda())
        dm = E.make_mask(det["delete"][0], torch.from_numpy(det["delete"][1]).cuda())
        if a == 2:
            d_dests[i] = (torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda"), torch.zeros((h, w), dtype=torch.float32, device="cuda"))
        d_ops.append((a, cm, dm, d_dests[i][0] if a == 2 else None, d_dests[i][1] if a == 2 else None))
    fr.ProcessSilhouettes(d_rgb, d_depth, d_ops)
    P._cmp("main depth", d_depth.cpu().numpy(), o_depth)
    for i in (0, 29):
        P._cmp(f"instance {i} depth", d_dests[i][1].cpu().numpy(), dests[i][1])
    assert (dests[29][1] == 0).all() and (dests[0][1] > 0).any()     # the last one copies from an already blanked frame


def test_compositing_equals_oracle():
    rng = np.random.default_rng(5)
    h, w = 375, 1242
    n = h * w
    L = H.oracle()

    def render(p_empty):
        d = rng.uniform(0.5, 30.0, (h, w)).astype(np.float32)
        d[rng.random((h, w)) < p_empty] = 0.0
        return rng.integers(0, 256, (h, w, 4), dtype=np.uint8), d

    bg_c, bg_d = render(0.2)
    inst = [render(0.7) for _ in range(19)]                                  # > 16 layers: two launches
    tints = [E.MATPLOTLIB2_PALETTE[i % len(E.MATPLOTLIB2_PALETTE)] for i in range(len(inst))]
    eng = _engine()
    fr = E.InstanceFrames(eng)
    for tint_strength, dim in ((1.0, 0.10), (0.0, -1.0), (0.35, 0.25)):
        want_c, want_d = bg_c.copy(), bg_d.copy()
        layers = (abi.InstanceLayer * len(inst))()
        for k, (c, d) in enumerate(inst):
            layers[k].d_color, layers[k].d_depth = c.ctypes.data, d.ctypes.data
            layers[k].tint = (C.c_int32 * 4)(*tints[k])
        L.oracle_composite_instances(H.vptr(want_c), H.vptr(want_d), n, layers, len(inst), dim, tint_strength)
        got_c, got_d = torch.from_numpy(bg_c).cuda(), torch.from_numpy(bg_d).cuda()
        d_layers = [(torch.from_numpy(c).cuda(), torch.from_numpy(d).cuda(), t) for (c, d), t in zip(inst, tints)]
        fr.CompositeInstances(got_c, got_d, d_layers, dim_factor=dim, tint_strength=tint_strength)
        P._cmp("composite colour", got_c.cpu().numpy(), want_c)
        P._cmp("composite depth", got_d.cpu().numpy(), want_d)
    # the two stand-alone entry points
    want_c, want_d = bg_c.copy(), bg_d.copy()
    L.oracle_composite_color(H.vptr(want_c), H.vptr(want_d), H.vptr(inst[0][0]), H.vptr(inst[0][1]), n, (C.c_int32 * 4)(*tints[3]), 0.6)
    got_c, got_d = torch.from_numpy(bg_c).cuda(), torch.from_numpy(bg_d).cuda()
    fr.CompositeColor(got_c, got_d, torch.from_numpy(inst[0][0]).cuda(), torch.from_numpy(inst[0][1]).cuda(), tints[3], 0.6)
    P._cmp("CompositeColor colour", got_c.cpu().numpy(), want_c)
    P._cmp("CompositeColor depth", got_d.cpu().numpy(), want_d)
    t, s = bg_d.copy(), inst[1][1].copy()
    L.oracle_composite_depth(H.vptr(t), H.vptr(s), n)
    g = torch.from_numpy(bg_d).cuda()
    fr.CompositeDepth(g, torch.from_numpy(inst[1][1]).cuda())
    P._cmp("CompositeDepth", g.cpu().numpy(), t)
