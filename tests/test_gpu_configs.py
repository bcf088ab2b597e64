"""Parity at the other BASELINE.json configurations (configs[2..4]) at sizes the oracle finishes in
seconds, plus size-independent properties of the decay sweep at a larger size."""
import ctypes as C

import numpy as np
import pytest
import torch

from dynslam_b200 import abi, engine as E, synth
from tests import hostlib as H
from tests import parity as P

pytestmark = pytest.mark.gpu


def test_config3_instance_volumes_masked_frames():
    """configs[2]: per-car volumes (mu 1.0, voxel 0.035, 7142 blocks; InstanceReconstructor.cpp:365-389) fed the
    frame masked to the car's silhouette (depth 0 / RGB 255 outside, :91-127) with the object pose — at the full
    1242x375 resolution and table size of the configuration."""
    scale = 1.0
    cfg = P.Cfg(scale=scale, numBlocks=7142, numBuckets=0x100000, excessSize=0x80000, voxelSize=0.035, mu=1.0, maxW=50,
                decay=(1, 2))
    scene = synth.StreetScene(seed=3, length_m=80.0)
    cars = [synth.MovingCar(i, seed=3) for i in range(2)]
    pairs = [P.Pair(cfg) for _ in cars]
    proj = synth.kitti_intrinsics() * np.float32(scale)
    w, h = pairs[0].w, pairs[0].h
    seen = [0, 0]
    for f in range(5):
        M = synth.kitti_pose(f)
        depth, rgb, ident = scene.render(M, w, h, float(proj[0]), float(proj[1]), float(proj[2]), float(proj[3]),
                                         extra_boxes=[c.box(f) for c in cars], want_ids=True)
        for i, car in enumerate(cars):
            d, c = synth.instance_frame(depth, rgb, ident, i)
            seen[i] += int((d > 0).sum())
            pairs[i].step(d, c, car.object_pose(f, M), proj, f)
    assert min(seen) > 500
    assert all(p.scene.lastFreeBlockId < cfg.numBlocks - 50 for p in pairs)
    # an instance that left the image: empty frame, nothing to integrate (Reco_CUDA.cu:372-378)
    d0 = np.zeros((h, w), dtype=np.float32)
    c0 = np.full((h, w, 4), 255, dtype=np.uint8)
    pairs[0].step(d0, c0, cars[0].object_pose(6, synth.kitti_pose(6)), proj, 6)


def test_config5_high_resolution_4mm():
    """configs[4]: 4 mm voxels, mu 16 mm: many small blocks, long chains with a small table."""
    cfg = P.Cfg(scale=0.25, frames=3, frame_step=1, numBlocks=131072, numBuckets=0x8000, excessSize=0x20000, voxelSize=0.004,
                mu=0.016, zmax=7.5, raycast=True)
    pair, _ = P.run_sequence(cfg)
    assert pair.rs.noVisibleBlocks > 5000
    assert cfg.excessSize - 1 - pair.scene.lastFreeExcessListId > 500


def _build_hires(numBlocks, frames, scale=0.5, numBuckets=0x40000, excessSize=0x80000):
    p = E.SceneParams(voxelSize=0.004, mu=0.016, maxW=50)
    scene = E.Scene(p, numBlocks, numBuckets, excessSize)
    w, h = int(round(synth.KITTI_W * scale)), int(round(synth.KITTI_H * scale))
    eng = E.Engine(scene, (w, h))
    reco = E.SceneReconstructionEngine(eng)
    vis = E.VisualisationEngine(eng, scene)
    rs = vis.CreateRenderState((w, h))
    reco.ResetScene(scene)
    street = synth.StreetScene(seed=4, length_m=60.0)
    for f in range(frames):
        depth, rgb, M, proj = synth.kitti_frame(street, f, scale=scale, zmax=8.0)
        v = E.View(torch.from_numpy(depth).cuda(), torch.from_numpy(rgb).cuda(), M, proj)
        reco.AllocateSceneFromDepth(scene, v, rs)
        reco.IntegrateIntoScene(scene, v, rs)
    return scene, eng, reco, rs


def test_config4_full_decay_properties_at_scale():
    """configs[3]: Decay(forceAllVoxels) over a few hundred thousand allocated blocks. Size-independent properties:
    conservation (freed + still allocated == allocated before), the free list stays a permutation, idempotence,
    and with maxWeight >= maxW everything is reclaimed and no entry points into the VBA any more."""
    numBlocks = 400000
    scene, eng, reco, rs = _build_hires(numBlocks, frames=6)
    allocated = numBlocks - 1 - scene.lastFreeBlockId
    assert allocated > 40000
    reco.Decay(scene, rs, 1, 0, True)                      # noisy voxels only
    freed1 = reco.GetDecayedBlockCount()
    st = scene.to_host()
    still = int((st["hash"]["ptr"] >= 0).sum())
    assert freed1 > 0 and freed1 + still == allocated
    assert scene.lastFreeBlockId == numBlocks - 1 - still
    free = st["allocationList"][:scene.lastFreeBlockId + 1]
    used = st["hash"]["ptr"][st["hash"]["ptr"] >= 0]
    assert len(np.unique(np.concatenate([free, used]))) == numBlocks      # a permutation: nothing lost, nothing doubled
    reco.Decay(scene, rs, 1, 0, True)                      # idempotent
    assert reco.GetDecayedBlockCount() == freed1
    reco.Decay(scene, rs, 255, 0, True)                    # everything is noise now
    st = scene.to_host()
    assert scene.lastFreeBlockId == numBlocks - 1 and not (st["hash"]["ptr"] >= 0).any()
    assert np.array_equal(np.sort(st["allocationList"]), np.arange(numBlocks, dtype=np.int32))
    vox = st["voxels"]
    assert (vox["w_depth"] == 0).all() and (vox["sdf"] == 32767).all()
    assert reco.GetDecayedBlockCount() == allocated


def test_config4_full_decay_matches_oracle():
    cfg = P.Cfg(scale=0.25, frames=3, frame_step=1, numBlocks=131072, numBuckets=0x8000, excessSize=0x20000, voxelSize=0.004,
                mu=0.016, zmax=7.5, raycast=False)
    pair, _ = P.run_sequence(cfg)
    pair.reco.Decay(pair.scene, pair.rs, 1, 1, True)
    freed = pair.L.oracle_decay(pair.host.engine, C.byref(pair.host.scene), C.byref(pair.host.rs), 1, 1, 1)
    assert freed > 100
    pair.compare_scene("full decay hires")
