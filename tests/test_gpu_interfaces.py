"""Parity of the remaining interface entry points: ForwardRender, CreatePointCloud and the
swapping engine. They are dead under DynSLAM's settings (useApproximateRaycast=false, TRACKER_EXTERNAL, swapping off)
but part of the engine interfaces, so they are implemented and checked bit for bit like the live ones."""
import ctypes as C

import numpy as np
import pytest
import torch

from dynslam_b200 import abi, engine as E, synth
from tests import hostlib as H
from tests import parity as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def built():
    cfg = P.Cfg(frames=4)
    pair, (gv, hv) = P.run_sequence(cfg)
    return cfg, pair, gv, hv


def test_forward_render(built):
    cfg, pair, gv, hv = built
    L = pair.L
    # a new pose: the previous raycast result is forward-projected into it, holes are re-cast
    M = synth.kitti_pose((cfg.frames - 1) * cfg.frame_step + 1)
    proj = synth.kitti_intrinsics() * np.float32(cfg.scale)
    depth, rgb = np.asarray(hv._keep[0]), np.asarray(hv._keep[1])
    g2, h2 = pair.views(depth, rgb, M, proj)
    frs = pair.vis.CreateRenderState((pair.w, pair.h), forward=True)
    # same inputs on both sides: visible list, min/max image and ray points of the live state
    for name in ("visibleBlockPositions", "entriesVisibleType", "renderingRangeImage", "raycastResult"):
        getattr(frs, name).copy_(getattr(pair.rs, name))
    frs.c.noVisibleBlocks = pair.rs.c.noVisibleBlocks
    n_gpu = pair.vis.ForwardRender(g2, frs)
    L.oracle_forward_render(C.byref(pair.host.scene), C.byref(pair.host.rs), C.byref(h2))
    assert n_gpu == pair.host.rs.noFwdProjMissingPoints > 0
    P._cmp("forwardProjection", frs.forwardProjection.cpu().numpy(), pair.host.forwardProjection.reshape(-1))
    P._cmp("missing list", frs.fwdProjMissingPoints.cpu().numpy()[:n_gpu], pair.host.fwdMissing[:n_gpu])
    P._cmp("forward image", frs.raycastImage.cpu().numpy(), pair.host.raycastImage.reshape(-1))


def test_point_cloud(built):
    cfg, pair, gv, hv = built
    L = pair.L
    dev = pair.scene.device
    n_px = pair.w * pair.h
    for skip in (0, 1):
        loc = torch.zeros(n_px * 4, dtype=torch.float32, device=dev)
        col = torch.zeros(n_px * 4, dtype=torch.float32, device=dev)
        n = pair.vis.CreatePointCloud(gv, pair.rs, loc, col, skipPoints=bool(skip))
        hl = np.zeros((n_px, 4), dtype=np.float32)
        hc = np.zeros((n_px, 4), dtype=np.float32)
        eye = abi.mat_to_c(np.eye(4, dtype=np.float32))
        hn = L.oracle_point_cloud(C.byref(pair.host.scene), C.byref(pair.host.rs), C.byref(hv), eye, skip, H.vptr(hl), H.vptr(hc))
        assert n == hn > 100
        P._cmp("pc rays", pair.rs.raycastResult.cpu().numpy(), pair.host.raycastResult.reshape(-1))
        P._cmp("pc image", pair.rs.raycastImage.cpu().numpy(), pair.host.raycastImage.reshape(-1))
        P._cmp("pc locations", loc.cpu().numpy().reshape(-1, 4)[:n], hl[:n])
        P._cmp("pc colours", col.cpu().numpy().reshape(-1, 4)[:n], hc[:n])


def test_swapping_out_and_in():
    """SaveToGlobalMemory / IntegrateGlobalIntoLocal (Swap_CUDA.cu:44-216): blocks whose state is 2 and that are not visible
    are moved to the transfer buffer, reset and freed (ptr -1); entries flagged 1 are merged back from the host store."""
    cfg = P.Cfg(frames=3, raycast=False, numBlocks=8192, numBuckets=0x2000, excessSize=0x1000)
    pair, _ = P.run_sequence(cfg)
    L = pair.L
    dev = pair.scene.device
    n_tot = cfg.numBuckets + cfg.excessSize
    st = pair.scene.to_host()
    used = np.nonzero(st["hash"]["ptr"] >= 0)[0]
    rng = np.random.RandomState(5)
    # swap states: every allocated entry "most recent in active memory" (2); make half of them invisible
    states = np.zeros(n_tot, dtype=np.uint8)
    states[used] = 2
    vis = pair.rs.entriesVisibleType.cpu().numpy().copy()
    hide = rng.choice(used, size=len(used) // 2, replace=False)
    vis[hide] = 0
    pair.rs.entriesVisibleType.copy_(torch.from_numpy(vis).to(dev))
    pair.host.visType[:] = vis
    # attach swap state arrays to both scenes
    d_states = torch.from_numpy(states.copy()).to(dev)
    pair.scene.swapStates = d_states
    pair.scene.c.d_swapStates = d_states.data_ptr()
    h_states = states.copy()
    pair.host.scene.d_swapStates = h_states.ctypes.data
    cache = E.GlobalCache(pair.scene)
    swap = E.SwappingEngine(pair.eng)
    # ---- swap out
    n_out = swap.SaveToGlobalMemory(pair.scene, pair.rs, cache)
    T = abi.TRANSFER_BLOCK_NUM
    h_sync = np.zeros(T * 512, dtype=abi.VOXEL_DTYPE)
    h_has = np.zeros(T, dtype=np.uint8)
    h_ids = np.zeros(T, dtype=np.int32)
    n_ref = L.oracle_swap_out(C.byref(pair.host.scene), C.byref(pair.host.rs), H.vptr(h_sync), H.vptr(h_has), H.vptr(h_ids))
    assert n_out == n_ref > 50
    P._cmp("needed ids (out)", cache.neededEntryIDs.cpu().numpy()[:n_out], h_ids[:n_out])
    P._cmp("synced blocks", cache.syncedVoxelBlocks.cpu().numpy()[:n_out * 4096], h_sync[:n_out * 512].view(np.uint8))
    P._cmp("swap states (out)", d_states.cpu().numpy(), h_states)
    pair.compare_scene("after swap out")
    # ---- swap in: flag the swapped-out entries that still have a VBA block?  The reference re-allocates them in
    # AllocateSceneFromDepth; here the merge itself is checked on entries that are resident: flag some resident ones 1.
    st2 = pair.scene.to_host()
    resident = np.nonzero(st2["hash"]["ptr"] >= 0)[0]
    pick = rng.choice(resident, size=min(200, len(resident)), replace=False)
    s_now = d_states.cpu().numpy().copy()
    s_now[pick] = 1
    d_states.copy_(torch.from_numpy(s_now).to(dev))
    h_states[:] = s_now
    # host store: give every picked entry a stored block (reuse blocks swapped out above, cyclically)
    stored_ids = sorted(cache.stored.keys())
    for k, entry in enumerate(sorted(int(x) for x in pick)):
        cache.stored[entry] = cache.stored[stored_ids[k % len(stored_ids)]].copy()
    n_in = swap.IntegrateGlobalIntoLocal(pair.scene, cache)
    ids = np.zeros(T, dtype=np.int32)
    n_in_ref = L.oracle_swap_list_in(C.byref(pair.host.scene), H.vptr(ids))
    assert n_in == n_in_ref == len(pick)
    blocks = np.zeros((n_in, 4096), dtype=np.uint8)
    for i in range(n_in):
        blocks[i] = cache.stored[int(ids[i])]
    L.oracle_swap_integrate_in(C.byref(pair.host.scene), H.vptr(blocks), H.vptr(ids), n_in)
    P._cmp("needed ids (in)", cache.neededEntryIDs.cpu().numpy()[:n_in], ids[:n_in])
    P._cmp("swap states (in)", d_states.cpu().numpy(), h_states)
    pair.compare_scene("after swap in")
