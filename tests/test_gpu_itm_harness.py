"""The drop-in, exercised through the reference's own C++ objects and virtual interfaces
(oracle/itm_harness.cpp): (1) the B200 shim classes behind ITMSceneReconstructionEngine /
IITMVisualisationEngine produce exactly the oracle's state; (2) the UNMODIFIED reference CUDA engines,
built for sm_100a, agree with it on the order-free invariants (the reference is nondeterministic and
compiled with --use_fast_math, SURVEY finding 4, so bit-exactness is not defined against it)."""
import ctypes as C

import numpy as np
import pytest

from dynslam_b200 import abi, synth
from tests import harnesslib as HL
from tests import hostlib as H

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not HL.available(), reason="oracle/_ref/libitmharness.so not built")]

NB, NE = 0x100000, 0x80000
SCALE, NUM_BLOCKS, FRAMES = 0.25, 32768, 6


def _frames():
    scene = synth.StreetScene(seed=6, length_m=80.0)
    return [synth.kitti_frame(scene, f * 2, scale=SCALE) for f in range(FRAMES)]


def _oracle_run(frames, decay):
    L = H.oracle()
    w, h = frames[0][0].shape[1], frames[0][0].shape[0]
    vol = H.HostVolume(NUM_BLOCKS, NB, NE, w, h)
    for depth, rgb, M, proj in frames:
        v = H.make_view(depth, rgb, M, proj)
        assert L.oracle_allocate_from_depth(vol.engine, C.byref(vol.scene), C.byref(vol.rs), C.byref(v), 0, 0) == 0
        L.oracle_integrate(vol.engine, C.byref(vol.scene), C.byref(vol.rs), C.byref(v), 0)
        L.oracle_expected_depths(C.byref(vol.scene), C.byref(vol.rs), C.byref(H.make_camera(M, proj)))
        L.oracle_icp_maps(C.byref(vol.scene), C.byref(vol.rs), C.byref(v), H.vptr(vol.points), H.vptr(vol.normals), 0)
        L.oracle_decay(vol.engine, C.byref(vol.scene), C.byref(vol.rs), decay[0], decay[1], 0)
    return vol


def test_shim_through_itmlib_objects_is_bit_exact():
    frames = _frames()
    decay = (2, 2)
    w, h = frames[0][0].shape[1], frames[0][0].shape[0]
    hs = HL.Harness(HL.B200_SHIM, w, h, frames[0][3], numBlocks=NUM_BLOCKS)
    for depth, rgb, M, proj in frames:
        hs.process_frame(depth, rgb, M, decay=decay)
    got, ctr = hs.download(), hs.counters()
    hs.close()
    vol = _oracle_run(frames, decay)
    L = H.oracle()
    assert ctr["lastFreeBlockId"] == vol.scene.lastFreeBlockId and ctr["noVisibleBlocks"] == vol.rs.noVisibleBlocks
    assert ctr["decayed"] == L.oracle_decayed_block_count(vol.engine) > 0
    for f in ("pos", "offset", "ptr", "allocatedTime"):
        assert np.array_equal(got["hash"][f], vol.hash[f]), f
    assert got["voxels"].tobytes() == vol.voxels.tobytes()
    assert got["rays"].tobytes() == vol.raycastResult.tobytes()
    assert got["image"].tobytes() == vol.raycastImage.tobytes()


def _by_pos(state):
    hs = state["hash"]
    used = np.nonzero(hs["ptr"] >= 0)[0]
    return {tuple(int(c) for c in hs["pos"][i]): int(hs["ptr"][i]) for i in used}


def test_reference_cuda_build_agrees_on_order_free_invariants():
    frames = _frames()
    w, h = frames[0][0].shape[1], frames[0][0].shape[0]
    res = {}
    for impl in (HL.REFERENCE_CUDA, HL.B200_SHIM):
        hs = HL.Harness(impl, w, h, frames[0][3], numBlocks=NUM_BLOCKS)
        for depth, rgb, M, proj in frames:
            hs.process_frame(depth, rgb, M, decay=None)
        res[impl] = (hs.download(), hs.counters())
        hs.close()
    (ref, rc), (own, oc) = res[HL.REFERENCE_CUDA], res[HL.B200_SHIM]
    pr, po = _by_pos(ref), _by_pos(own)
    common = set(pr) & set(po)
    # same set of allocated block positions, up to same-frame bucket races / skipped contended steps in the reference
    assert len(common) >= 0.995 * max(len(pr), len(po)), (len(pr), len(po), len(common))
    assert abs(rc["noVisibleBlocks"] - oc["noVisibleBlocks"]) <= 0.01 * oc["noVisibleBlocks"] + 2
    # per-position voxel contents: weights identical, TSDF within 1e-5 (float) where the nearest-pixel lookup agrees;
    # fast-math projection flips a few lookups, so allow a small fraction of outliers and report it
    rng = np.random.RandomState(0)
    pick = [sorted(common)[i] for i in rng.choice(len(common), size=min(1500, len(common)), replace=False)]
    n = bad_w = bad_sdf = 0
    for p in pick:
        a = ref["voxels"][pr[p] * 512:(pr[p] + 1) * 512]
        b = own["voxels"][po[p] * 512:(po[p] + 1) * 512]
        n += 512
        bad_w += int((a["w_depth"] != b["w_depth"]).sum())
        d = np.abs(a["sdf"].astype(np.float32) / 32767.0 - b["sdf"].astype(np.float32) / 32767.0)
        bad_sdf += int((d > 1e-5 + 1.0 / 32767.0).sum())     # 1 LSB of the short quantisation + 1e-5
    assert bad_w / n < 0.03 and bad_sdf / n < 0.03, (bad_w / n, bad_sdf / n)   # measured on B200: ~1.3 % / ~1.1 %
    # the raycast images agree on almost every pixel
    found_r, found_o = ref["rays"][..., 3] > 0, own["rays"][..., 3] > 0
    assert (found_r == found_o).mean() > 0.98
    both = found_r & found_o
    assert np.abs(ref["rays"][both][:, :3] - own["rays"][both][:, :3]).max(axis=1).mean() < 0.05   # voxel units


def test_view_builder_through_itmlib_objects():
    """ITMViewBuilder_B200 behind the abstract ITMViewBuilder, called the way ITMMainEngine::ProcessFrame calls it
    (host ITMUChar4Image / ITMShortImage in): close to the oracle (libm vs CUDA exp), zero border exactly; the
    reference's ITMViewBuilder_CUDA (fast-math exp and division) agrees within 1e-4 relative."""
    from tests import viewlib
    raw, rgb = viewlib.raw_kitti_frame(scale=0.5)
    h, w = raw.shape
    proj = (353.5, 353.5, w / 2.0, h / 2.0)
    L = H.oracle()
    calib = abi.ViewCalib()
    calib.trafoType, calib.useBilateralFilter = 1, 1
    calib.params = (C.c_float * 2)(1.0 / 1000.0, 0.0)
    want, scratch = np.zeros((h, w), np.float32), np.zeros((h, w), np.float32)
    L.oracle_update_view(H.vptr(raw), w, h, C.byref(calib), H.vptr(want), H.vptr(scratch), None, None)
    got = {}
    for impl in (HL.B200_SHIM, HL.REFERENCE_CUDA):
        vb = HL.ViewBuilderHarness(impl, w, h, proj)
        first = vb.update_view(raw, rgb)
        again = vb.update_view(raw, rgb)                    # second frame through the same builder: same result
        assert np.array_equal(first, again)
        got[impl] = first
        vb.close()
    for impl, tol in ((HL.B200_SHIM, 5e-6), (HL.REFERENCE_CUDA, 1e-4)):
        g = got[impl]
        assert ((g == -1.0) == (want == -1.0)).all()
        err = np.abs(g.astype(np.float64) - want) / np.maximum(1.0, np.abs(want))
        assert err.max() <= tol, (impl, err.max())
        assert (g[:2] == 0).all() and (g[-2:] == 0).all() and (g[:, :2] == 0).all() and (g[:, -2:] == 0).all()


@pytest.mark.skipif(not HL.patched_available(), reason="oracle/_ref/libitmpatched.so not built (integration/build_patched.sh)")
def test_patched_main_engine_runs_both_backends():
    """The binding a maintainer adds, compiled and RUN: integration/itmlib_b200.patch applied to the reference's ITMLib, the
    whole library rebuilt, and the reference's own top-level object — ITMMainEngine, the base class of DynSLAM's
    InfiniTamDriver — driven frame by frame with settings->engineBackend = BACKEND_B200 and BACKEND_REFERENCE. Everything
    around the engines (view building, ITMDenseMapper::ProcessFrame, ITMTrackingController::Prepare, GetImage) is the
    reference's host code. The two back-ends are compared on order-free invariants (the reference CUDA build is
    nondeterministic and uses --use_fast_math)."""
    scale = 0.5
    w, h = int(round(synth.KITTI_W * scale)), int(round(synth.KITTI_H * scale))
    street = synth.StreetScene(seed=6, length_m=60.0)
    frames = [synth.kitti_frame(street, f, scale=scale) for f in range(6)]
    res = {}
    for backend in (0, 1):
        eng = HL.PatchedMainEngine(backend, w, h, frames[0][3], numBlocks=65536)
        for depth, rgb, M, proj in frames:
            eng.process_frame(np.round(depth * 1000.0).astype(np.int16), rgb, M)
        res[backend] = (eng.counters(), eng.raycast_image())
        eng.close()
    (c0, img0), (c1, img1) = res[0], res[1]
    assert c1["allocatedEntries"] > 2000
    assert abs(c1["allocatedEntries"] - c0["allocatedEntries"]) <= 0.01 * c0["allocatedEntries"]
    assert c1["lastFreeBlockId"] == 65536 - 1 - c1["allocatedEntries"]
    hit0, hit1 = img0[..., 0] > 0, img1[..., 0] > 0
    assert hit1.mean() > 0.3 and abs(hit0.mean() - hit1.mean()) < 0.02
    both = hit0 & hit1
    # shaded grey levels at the pixels both back-ends hit: the shading takes normals from TSDF differences, the reference build
    # computes them with --use_fast_math on a volume whose weight-1 voxels (the ones Decay(1, 3) removes) depend on its
    # nondeterministic allocation order — so the bulk must agree closely, a tail may differ
    d = np.abs(img0[..., 0].astype(np.int32) - img1[..., 0].astype(np.int32))[both]
    stats = dict(mean=float(d.mean()), median=float(np.median(d)), p90=float(np.percentile(d, 90)), over16=float((d > 16).mean()))
    assert stats["median"] <= 2.0 and stats["over16"] < 0.10, stats
