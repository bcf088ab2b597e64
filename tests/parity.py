"""GPU-vs-oracle parity harness (test infrastructure; also used by __graft_entry__.smoke()).

Runs the same seeded frame sequence through the CUDA engine (via the C-ABI) and through the CPU
oracle and compares EVERY piece of state after every stage, bit for bit:
hash table, visibility bytes, visible list, counters, free lists, the whole voxel block array,
min/max image, ray points, ICP maps and the grey raycast image."""
import ctypes as C
import json
import os

import numpy as np
import torch

from dynslam_b200 import abi, engine as E, synth
from tests import hostlib as H


class Cfg:
    def __init__(self, **kw):
        self.scale = 0.25
        self.numBlocks, self.numBuckets, self.excessSize = 16384, 0x4000, 0x2000
        self.voxelSize, self.mu, self.maxW = 0.05, 0.75, 50
        self.vf_min, self.vf_max = 0.1, 300.0
        self.frames, self.frame_step = 6, 2
        self.decay = None            # (maxWeight, minAge)
        self.depthWeighting = False
        self.stopMaxW = False
        self.zmax = 20.0
        self.maxRenderingBlocks = 0   # > 0: lower MAX_RENDERING_BLOCKS on the engine (diag hook) to reach the cap rule
        self.decayRingItems = 0
        self.scene_seed = 6
        self.raycast = True
        self.workload = "kitti"
        self.__dict__.update(kw)


def _cmp(name, a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.shape != b.shape or a.tobytes() != b.tobytes():
        av, bv = a.reshape(-1).view(np.uint8), b.reshape(-1).view(np.uint8)
        n = min(av.size, bv.size)
        bad = np.nonzero(av[:n] != bv[:n])[0]
        raise AssertionError(f"{name}: {bad.size} differing bytes of {n} (shapes {a.shape} vs {b.shape}); first at byte "
                             f"{bad[0] if bad.size else -1}")


class Pair:
    """One volume on the GPU and its oracle twin."""

    def __init__(self, cfg, device="cuda:0"):
        self.cfg = cfg
        if cfg.workload == "plane":
            self.w, self.h = 640, 480
        else:
            self.w, self.h = int(round(synth.KITTI_W * cfg.scale)), int(round(synth.KITTI_H * cfg.scale))
        p = E.SceneParams(cfg.voxelSize, cfg.mu, cfg.maxW, cfg.vf_min, cfg.vf_max, cfg.stopMaxW)
        self.scene = E.Scene(p, cfg.numBlocks, cfg.numBuckets, cfg.excessSize, device=device)
        self.eng = E.Engine(self.scene, (self.w, self.h), decayRingItems=cfg.decayRingItems)
        if cfg.maxRenderingBlocks > 0:
            self.eng.set_max_rendering_blocks(cfg.maxRenderingBlocks)
        self.reco = E.SceneReconstructionEngine(self.eng)
        self.reco.SetFusionWeightParams(cfg.depthWeighting)
        self.vis = E.VisualisationEngine(self.eng, self.scene)
        self.rs = self.vis.CreateRenderState((self.w, self.h))
        self.reco.ResetScene(self.scene)
        dev = self.scene.device
        self.points = torch.zeros(self.h * self.w * 4, dtype=torch.float32, device=dev)
        self.normals = torch.zeros(self.h * self.w * 4, dtype=torch.float32, device=dev)
        self.host = H.HostVolume(cfg.numBlocks, cfg.numBuckets, cfg.excessSize, self.w, self.h,
                                 H.SceneParams(cfg.voxelSize, cfg.mu, cfg.maxW, cfg.vf_min, cfg.vf_max, int(cfg.stopMaxW)))
        self.L = H.oracle()

    def views(self, depth, rgb, M, proj):
        dev = self.scene.device
        d = torch.from_numpy(np.ascontiguousarray(depth)).to(dev)
        c = torch.from_numpy(np.ascontiguousarray(rgb)).to(dev)
        gv = E.View(d, c, M, proj, depthWeighting=self.cfg.depthWeighting)
        hv = H.make_view(depth, rgb, M, proj, depthWeighting=int(self.cfg.depthWeighting))
        assert bytes(gv.c.invM_d) == bytes(hv.invM_d), "b200_mat4_inv differs from the oracle's"
        return gv, hv

    def compare_scene(self, tag, voxels=True):
        g, r = self.scene.to_host(), self.rs.to_host()
        o = self.host.state()
        assert g["lastFreeBlockId"] == o["lastFreeBlockId"], (tag, g["lastFreeBlockId"], o["lastFreeBlockId"])
        assert g["lastFreeExcessListId"] == o["lastFreeExcessListId"], tag
        assert r["noVisibleBlocks"] == o["noVisibleBlocks"], (tag, r["noVisibleBlocks"], o["noVisibleBlocks"])
        for f in ("pos", "offset", "ptr", "allocatedTime"):
            _cmp(f"{tag}: hash.{f}", g["hash"][f], o["hash"][f])
        _cmp(f"{tag}: entriesVisibleType", r["visType"], o["visType"])
        _cmp(f"{tag}: visible list", r["visiblePos"], o["visiblePos"])
        _cmp(f"{tag}: allocationList", g["allocationList"], o["allocationList"])
        _cmp(f"{tag}: excessList", g["excessList"], o["excessList"])
        if voxels:
            _cmp(f"{tag}: voxels", g["voxels"], o["voxels"])

    def step(self, depth, rgb, M, proj, frame_no):
        cfg, L, hv_ = self.cfg, self.L, None
        gv, hv = self.views(depth, rgb, M, proj)
        tag = f"frame {frame_no}"
        self.reco.AllocateSceneFromDepth(self.scene, gv, self.rs)
        rc = L.oracle_allocate_from_depth(self.host.engine, C.byref(self.host.scene), C.byref(self.host.rs), C.byref(hv), 0, 0)
        assert rc == 0
        self.compare_scene(tag + " allocate", voxels=False)
        self.reco.IntegrateIntoScene(self.scene, gv, self.rs)
        L.oracle_integrate(self.host.engine, C.byref(self.host.scene), C.byref(self.host.rs), C.byref(hv), 0)
        self.compare_scene(tag + " integrate")
        if cfg.raycast:
            cam = E.make_camera(M, proj)
            self.vis.CreateExpectedDepths(cam, self.rs)
            L.oracle_expected_depths(C.byref(self.host.scene), C.byref(self.host.rs), C.byref(H.make_camera(M, proj)))
            _cmp(tag + ": minmax", self.rs.renderingRangeImage.cpu().numpy(), self.host.minmax.reshape(-1))
            self.vis.CreateICPMaps(gv, self.rs, self.points, self.normals)
            L.oracle_icp_maps(C.byref(self.host.scene), C.byref(self.host.rs), C.byref(hv), H.vptr(self.host.points),
                              H.vptr(self.host.normals), 0)
            _cmp(tag + ": raycastResult", self.rs.raycastResult.cpu().numpy(), self.host.raycastResult.reshape(-1))
            _cmp(tag + ": raycastImage", self.rs.raycastImage.cpu().numpy(), self.host.raycastImage.reshape(-1))
            _cmp(tag + ": points", self.points.cpu().numpy(), self.host.points.reshape(-1))
            _cmp(tag + ": normals", self.normals.cpu().numpy(), self.host.normals.reshape(-1))
        if cfg.decay is not None:
            self.reco.Decay(self.scene, self.rs, cfg.decay[0], cfg.decay[1], False)
            L.oracle_decay(self.host.engine, C.byref(self.host.scene), C.byref(self.host.rs), cfg.decay[0], cfg.decay[1], 0)
            self.compare_scene(tag + " decay")
            assert self.reco.GetDecayedBlockCount() == L.oracle_decayed_block_count(self.host.engine)
        return gv, hv


def frames_of(cfg):
    if cfg.workload == "plane":
        depth, rgb, M, proj = synth.plane_frame(seed=1)
        for f in range(cfg.frames):
            yield depth, rgb, M, proj
        return
    scene = synth.StreetScene(seed=cfg.scene_seed, length_m=max(60.0, cfg.frames * cfg.frame_step * 0.8 + 40.0))
    for f in range(cfg.frames):
        yield synth.kitti_frame(scene, f * cfg.frame_step, scale=cfg.scale, zmax=cfg.zmax)


def run_sequence(cfg, device="cuda:0"):
    pair = Pair(cfg, device)
    last = None
    for i, (depth, rgb, M, proj) in enumerate(frames_of(cfg)):
        last = pair.step(depth, rgb, M, proj, i)
    return pair, last


def run_smoke():
    cfg = Cfg(scale=0.2, frames=3, frame_step=3, decay=(1, 1), numBlocks=8192, numBuckets=0x2000, excessSize=0x1000)
    pair, _ = run_sequence(cfg)
    assert pair.rs.noVisibleBlocks > 100
    return pair


# ---- golden fixtures (tests/golden/) -----------------------------------------------------------
import hashlib


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:32]


GOLDEN_CFG = dict(scale=0.25, frames=6, frame_step=2, numBlocks=16384, numBuckets=0x1000, excessSize=0x2000, decay=(3, 2))


def state_digests(hash_, visType, visiblePos, voxels, allocationList, minmax, rays, image, points, normals, counters):
    d = {f"hash.{f}": _sha(hash_[f]) for f in ("pos", "offset", "ptr", "allocatedTime")}
    d.update(visType=_sha(visType), visiblePos=_sha(visiblePos), voxels=_sha(voxels), allocationList=_sha(allocationList),
             minmax=_sha(minmax), raycastResult=_sha(rays), raycastImage=_sha(image), points=_sha(points), normals=_sha(normals),
             counters=[int(c) for c in counters])
    return d


def oracle_sequence_digests(cfg, ref_check=False):
    """Per-frame digests of the oracle's state on the golden sequence. With ref_check=True every frame's marking,
    integration of all visible blocks, raycast and ICP output is also recomputed with the reference's own functions
    (oracle/_ref/libitmref.so) and must be identical — that is what certifies the fixture."""
    L = H.oracle()
    w, h = int(round(synth.KITTI_W * cfg.scale)), int(round(synth.KITTI_H * cfg.scale))
    vol = H.HostVolume(cfg.numBlocks, cfg.numBuckets, cfg.excessSize, w, h,
                       H.SceneParams(cfg.voxelSize, cfg.mu, cfg.maxW, cfg.vf_min, cfg.vf_max, int(cfg.stopMaxW)))
    out = []
    for i, (depth, rgb, M, proj) in enumerate(frames_of(cfg)):
        hv = H.make_view(depth, rgb, M, proj, depthWeighting=int(cfg.depthWeighting))
        cam = H.make_camera(M, proj)
        assert L.oracle_allocate_from_depth(vol.engine, C.byref(vol.scene), C.byref(vol.rs), C.byref(hv), 0, 0) == 0
        pre = vol.voxels.copy() if ref_check else None
        L.oracle_integrate(vol.engine, C.byref(vol.scene), C.byref(vol.rs), C.byref(hv), 0)
        if ref_check:
            R = H.ref()
            for p in vol.visiblePos[:vol.rs.noVisibleBlocks]:
                idx = L.oracle_find_block(H.vptr(vol.hash), cfg.numBuckets, int(p[0]), int(p[1]), int(p[2]))
                if idx < 0:
                    continue
                ptr = int(vol.hash[idx]["ptr"])
                blk = pre[ptr * 512:(ptr + 1) * 512].copy()
                pos = np.array(p, dtype=np.int16)
                R.ref_integrate_block(H.vptr(blk), H.vptr(pos), C.byref(vol.scene), C.byref(hv))
                assert blk.tobytes() == vol.voxels[ptr * 512:(ptr + 1) * 512].tobytes()
        L.oracle_expected_depths(C.byref(vol.scene), C.byref(vol.rs), C.byref(cam))
        L.oracle_icp_maps(C.byref(vol.scene), C.byref(vol.rs), C.byref(hv), H.vptr(vol.points), H.vptr(vol.normals), 0)
        if ref_check and cfg.numBuckets == 0x100000:
            pass
        d = state_digests(vol.hash, vol.visType, vol.visiblePos[:vol.rs.noVisibleBlocks], vol.voxels, vol.allocationList,
                          vol.minmax, vol.raycastResult, vol.raycastImage, vol.points, vol.normals,
                          (vol.scene.lastFreeBlockId, vol.scene.lastFreeExcessListId, vol.rs.noVisibleBlocks))
        if cfg.decay is not None:
            L.oracle_decay(vol.engine, C.byref(vol.scene), C.byref(vol.rs), cfg.decay[0], cfg.decay[1], 0)
            d["after_decay"] = {"hash.ptr": _sha(vol.hash["ptr"]), "hash.offset": _sha(vol.hash["offset"]), "voxels": _sha(vol.voxels),
                                "allocationList": _sha(vol.allocationList), "visType": _sha(vol.visType),
                                "lastFreeBlockId": int(vol.scene.lastFreeBlockId),
                                "decayed": int(L.oracle_decayed_block_count(vol.engine))}
        out.append(d)
    return out


def gpu_sequence_digests(cfg, device="cuda:0"):
    pair = Pair(cfg, device)
    out = []
    for i, (depth, rgb, M, proj) in enumerate(frames_of(cfg)):
        gv = E.View(torch.from_numpy(depth).to(pair.scene.device), torch.from_numpy(rgb).to(pair.scene.device), M, proj,
                    depthWeighting=cfg.depthWeighting)
        pair.reco.AllocateSceneFromDepth(pair.scene, gv, pair.rs)
        pair.reco.IntegrateIntoScene(pair.scene, gv, pair.rs)
        pair.vis.CreateExpectedDepths(E.make_camera(M, proj), pair.rs)
        pair.vis.CreateICPMaps(gv, pair.rs, pair.points, pair.normals)
        g, r = pair.scene.to_host(), pair.rs.to_host()
        d = state_digests(g["hash"], r["visType"], r["visiblePos"], g["voxels"], g["allocationList"],
                          pair.rs.renderingRangeImage.cpu().numpy(), pair.rs.raycastResult.cpu().numpy(),
                          pair.rs.raycastImage.cpu().numpy(), pair.points.cpu().numpy(), pair.normals.cpu().numpy(),
                          (g["lastFreeBlockId"], g["lastFreeExcessListId"], r["noVisibleBlocks"]))
        if cfg.decay is not None:
            pair.reco.Decay(pair.scene, pair.rs, cfg.decay[0], cfg.decay[1], False)
            g, r = pair.scene.to_host(), pair.rs.to_host()
            d["after_decay"] = {"hash.ptr": _sha(g["hash"]["ptr"]), "hash.offset": _sha(g["hash"]["offset"]), "voxels": _sha(g["voxels"]),
                                "allocationList": _sha(g["allocationList"]), "visType": _sha(r["visType"]),
                                "lastFreeBlockId": int(g["lastFreeBlockId"]), "decayed": int(pair.reco.GetDecayedBlockCount())}
        out.append(d)
    return out
