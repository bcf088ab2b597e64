"""Parity at the sizes bench.py times: the headline configuration configs[1] at its own size —
1242x375, 0x100000 buckets + 0x80000 excess entries, 0x60000 voxel blocks — stage by stage against the oracle, a run long
enough for partial decay with the shipped min_decay_age 200 to fire, the >1024-candidate decay path and a wrapping
snapshot ring at that size, and the bounded decay queue. Bit-exact, like every other parity test."""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest
import torch

from dynslam_b200 import abi, engine as E, synth
from tests import hostlib as H
from tests import parity as P

pytestmark = pytest.mark.gpu

FULL = dict(scale=1.0, numBlocks=0x60000, numBuckets=0x100000, excessSize=0x80000, frame_step=1)


def test_headline_config_stage_by_stage():
    """configs[1] at full size, every stage of 6 consecutive frames compared (hash, free lists, visibility bytes, visible
    list, the whole 1.5 GiB voxel array, min/max image, rays, ICP maps, image), decay with a short minAge so it fires."""
    cfg = P.Cfg(frames=6, decay=(1, 2), **FULL)
    pair, _ = P.run_sequence(cfg)
    assert pair.rs.noVisibleBlocks > 4000 and pair.w == synth.KITTI_W and pair.h == synth.KITTI_H


def _frames(n, seed=6):
    street = synth.StreetScene(seed=seed, length_m=n * 0.8 + 60.0)
    with ThreadPoolExecutor(16) as ex:          # numpy releases the GIL in the render's array ops
        return list(ex.map(lambda f: synth.kitti_frame(street, f), range(n)))


def _oracle_frame(L, vol, hv, cam, decay):
    """the oracle's frame in its deterministic configuration: serial marking, OpenMP only where results cannot depend on it"""
    assert L.oracle_allocate_from_depth(vol.engine, C.byref(vol.scene), C.byref(vol.rs), C.byref(hv), 0, 0) == 0
    L.oracle_integrate(vol.engine, C.byref(vol.scene), C.byref(vol.rs), C.byref(hv), 1)
    L.oracle_expected_depths(C.byref(vol.scene), C.byref(vol.rs), C.byref(cam))
    L.oracle_icp_maps(C.byref(vol.scene), C.byref(vol.rs), C.byref(hv), H.vptr(vol.points), H.vptr(vol.normals), 1)
    if decay is not None:
        L.oracle_decay(vol.engine, C.byref(vol.scene), C.byref(vol.rs), decay[0], decay[1], 0)


def test_headline_config_long_run_default_decay():
    """The state bench.py times: the fused frame (b200_process_frame_async, no host synchronisation) over 206 full-size frames
    with the shipped decay parameters (maxWeight 1, minAge 200): the decay queue fills for 200 frames and partial decay runs on
    the last ones. Compared with the oracle after frames 3, 200, 203 and 206."""
    n, decay = 206, (1, 200)
    frames = _frames(n)
    cfg = P.Cfg(frames=n, decay=decay, **FULL)
    pair = P.Pair(cfg)
    L = pair.L
    dev = pair.scene.device
    checkpoints = {3, 200, 203, 206}
    freed_before = 0
    for i, (depth, rgb, M, proj) in enumerate(frames):
        gv = E.View(torch.from_numpy(depth).to(dev), torch.from_numpy(rgb).to(dev), M, proj)
        pair.eng.process_frame_async(pair.rs, gv, pair.points, pair.normals, decay=decay)
        _oracle_frame(L, pair.host, H.make_view(depth, rgb, M, proj), H.make_camera(M, proj), decay)
        if i + 1 in checkpoints:
            pair.eng.sync(pair.rs)
            tag = f"fused frame {i + 1}"
            pair.compare_scene(tag)
            P._cmp(tag + ": minmax", pair.rs.renderingRangeImage.cpu().numpy(), pair.host.minmax.reshape(-1))
            P._cmp(tag + ": raycastResult", pair.rs.raycastResult.cpu().numpy(), pair.host.raycastResult.reshape(-1))
            P._cmp(tag + ": raycastImage", pair.rs.raycastImage.cpu().numpy(), pair.host.raycastImage.reshape(-1))
            P._cmp(tag + ": points", pair.points.cpu().numpy(), pair.host.points.reshape(-1))
            assert pair.reco.GetDecayedBlockCount() == L.oracle_decayed_block_count(pair.host.engine)
            if i + 1 == 200:
                freed_before = pair.reco.GetDecayedBlockCount()
    assert freed_before == 0 and pair.reco.GetDecayedBlockCount() > 0     # decay really fired only after frame 200
    assert pair.scene.lastFreeBlockId < cfg.numBlocks - 50000             # ~10^5 blocks allocated, as in the bench


def test_headline_config_mass_deletion_and_ring_wrap():
    """Full size, decay (maxWeight 50, minAge 3): every block seen three frames ago is emptied and deleted — ~4.5 k claims
    per call, beyond k_decay_commit's 1024-candidate fast path — with a snapshot ring of 24 k items that the write cursor
    laps every five frames (positions are taken modulo the ring) while every live snapshot still fits."""
    cfg = P.Cfg(frames=9, decay=(50, 3), decayRingItems=24000, raycast=False, **FULL)
    pair, _ = P.run_sequence(cfg)
    assert pair.reco.GetDecayedBlockCount() > 4 * 1024
    assert pair.eng.stats().droppedSnapshots == 0


def test_decay_queue_and_ring_overflow_are_not_fatal():
    """the reference's queue of visible-list copies is unbounded; ours holds 4095 frames / decayRingItems items and
    drops the OLDEST snapshots beyond that — no error, the engine stays usable, nothing is lost or doubled."""
    # (a) more frames than queue slots without any Decay() call (--voxel_decay=false)
    cfg = P.Cfg(scale=0.1, numBlocks=16384, numBuckets=0x2000, excessSize=0x1000, raycast=False)
    pair = P.Pair(cfg)
    dev = pair.scene.device
    street = synth.StreetScene(seed=6, length_m=80.0)
    views = []
    for f in range(8):
        depth, rgb, M, proj = synth.kitti_frame(street, f, scale=cfg.scale)
        views.append(E.View(torch.from_numpy(depth).to(dev), torch.from_numpy(rgb).to(dev), M, proj))
    for i in range(4200):
        pair.eng.process_frame_async(pair.rs, views[i % 8], None, None, decay=None, raycast=False)
    pair.eng.sync(pair.rs)
    assert pair.eng.stats().droppedSnapshots >= 4200 - 4095
    for _ in range(3):                                   # Decay() still works on what the queue kept
        pair.reco.Decay(pair.scene, pair.rs, 1, 0, False)
    st = pair.scene.to_host()
    used = st["hash"]["ptr"][st["hash"]["ptr"] >= 0]
    free = st["allocationList"][:pair.scene.lastFreeBlockId + 1]
    assert len(np.unique(np.concatenate([free, used]))) == cfg.numBlocks
    # (b) a ring too small for min_decay_age snapshots: the overwritten ones are swept as empty
    cfg = P.Cfg(scale=0.25, frames=14, decay=(3, 6), decayRingItems=2500, raycast=False)
    pair = P.Pair(cfg)
    dev = pair.scene.device
    for depth, rgb, M, proj in P.frames_of(cfg):
        gv = E.View(torch.from_numpy(depth).to(dev), torch.from_numpy(rgb).to(dev), M, proj)
        pair.reco.AllocateSceneFromDepth(pair.scene, gv, pair.rs)        # synchronous calls: must not raise
        pair.reco.IntegrateIntoScene(pair.scene, gv, pair.rs)
        pair.reco.Decay(pair.scene, pair.rs, cfg.decay[0], cfg.decay[1], False)
    assert pair.rs.noVisibleBlocks > 1000 and pair.eng.stats().droppedSnapshots > 0
    st = pair.scene.to_host()
    used = st["hash"]["ptr"][st["hash"]["ptr"] >= 0]
    free = st["allocationList"][:pair.scene.lastFreeBlockId + 1]
    assert len(np.unique(np.concatenate([free, used]))) == cfg.numBlocks
    # ... and the engine keeps working after ResetScene (nothing sticky survives the overflow)
    pair.reco.ResetScene(pair.scene)
    assert pair.scene.lastFreeBlockId == cfg.numBlocks - 1
    for depth, rgb, M, proj in list(P.frames_of(cfg))[:2]:
        gv = E.View(torch.from_numpy(depth).to(dev), torch.from_numpy(rgb).to(dev), M, proj)
        pair.reco.AllocateSceneFromDepth(pair.scene, gv, pair.rs)
        pair.reco.IntegrateIntoScene(pair.scene, gv, pair.rs)
    assert pair.rs.noVisibleBlocks > 1000
