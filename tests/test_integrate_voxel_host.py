"""CPU test of the default IntegrateIntoScene kernel's per-voxel logic.

dynslam_b200/csrc/integrate_voxel.cuh is `__host__ __device__`; tests/hostcheck compiles those very functions for the host
and compares, voxel by voxel, the kernel's fast path (pose products -> v3_stage_a -> depth fetch -> v3_stage_b ->
done / v3_colour / generic path) with the generic per-voxel code that evaluates the reference's expressions with `/`
(DA/ITMSceneReconstructionEngine.h:14-171). Bit-exact, no tolerance. The GPU-only ingredient (MUFU.RCP inside rcp_nr) is
covered on the device by tests/test_gpu_parity.py::test_division_sequences_equal_ieee_division."""
import ctypes as C
import os
import subprocess

import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcheck")


@pytest.fixture(scope="module")
def lib():
    subprocess.run(["make", "-C", HERE, "all"], check=True, stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(HERE, "libhostcheck.so"))
    L.hostcheck_integrate.restype = C.c_longlong
    L.hostcheck_integrate.argtypes = [C.c_longlong, C.c_uint, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    L.hostcheck_integrate_v4.restype = C.c_longlong
    L.hostcheck_integrate_v4.argtypes = [C.c_longlong, C.c_uint, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    return L


# (mu, voxelSize, depthWeighting, maxW, identityPose)
CASES = [
    (0.75, 0.05, 0, 50, 0),       # DynSLAM defaults (ITMLibSettings.cpp:23)
    (0.75, 0.05, 0, 50, 1),       # identity pose: exact zeros in the camera coordinates -> generic path
    (1.0, 0.035, 1, 50, 0),       # instance volumes (InstanceReconstructor.cpp:372-379) + depth weighting
    (0.016, 0.004, 0, 50, 0),     # configs[4], 4 mm voxels
    (4.0, 0.05, 0, 100, 0),       # mu >= 4: rejected voxels pass the colour gate (SURVEY 8a')
    (0.032, 0.008, 1, 3, 1),      # configs[0] + early weight saturation
    (0.75, 0.05, 0, 50000, 0),    # --dynamic_weights: maxW 50000, stored weight wraps (DS/DynSLAMGUI.cpp:1217-1219)
]


@pytest.mark.parametrize("mu,voxel,dw,maxw,ident", CASES)
def test_fast_path_equals_generic_path(lib, mu, voxel, dw, maxw, ident):
    st = (C.c_longlong * 6)()
    bad = 0
    for seed in (1, 8):
        bad += lib.hostcheck_integrate(1200, seed, mu, voxel, dw, maxw, ident, st)
    voxels, fast, colour, generic, changed, behind = list(st)
    assert bad == 0
    assert voxels == 2 * 1200 * 512 and fast + generic == voxels
    assert changed > voxels // 20 and behind > voxels // 20 and colour > 0   # the interesting branches were exercised
    if mu >= 4.0 or ident:
        assert generic > 1000
    else:
        assert generic < voxels // 1000   # behind-the-camera voxels must not be sent to the generic path


@pytest.mark.parametrize("mu,voxel,dw,maxw,ident", CASES)
def test_v4_pair_path_equals_generic_path(lib, mu, voxel, dw, maxw, ident):
    """The default kernel's arithmetic (V4: v4_stage_a / v4_stage_b on pairs of x-adjacent voxels — packed FADD2 / FMUL2 / FFMA2
    on the device, the same operations element by element here) against the generic per-voxel code. Bit-exact."""
    st = (C.c_longlong * 8)()
    bad = 0
    for seed in (2, 9):
        bad += lib.hostcheck_integrate_v4(1200, seed, mu, voxel, dw, maxw, ident, 0, st)
    voxels, fast, colour, generic, changed = list(st)[:5]
    assert bad == 0
    assert voxels == 2 * 1200 * 512 and fast + generic == voxels
    assert changed > voxels // 20 and colour > 0


@pytest.mark.parametrize("mu,voxel,dw,maxw,ident", [c for c in CASES if c[0] < 4.0])
def test_v4_tolerance_mode_is_within_one_lsb(lib, mu, voxel, dw, maxw, ident):
    """B200_INTEGRATE_IMPL=fast (MUFU reciprocals, quotients as products, contracted sums): weights and the update / skip
    decision stay exact; a TSDF code moves by at most 1 LSB (3.05e-5 after SDF_valueToFloat), on a small share of the updated
    voxels; the colour gate |eta / mu| <= 0.25 may flip on a tie. (On the host the reciprocal is IEEE's, so the device shows
    slightly more flips: tests/test_gpu_parity.py::test_integrate_tolerance_mode counts them there.)"""
    st = (C.c_longlong * 8)()
    bad = 0
    for seed in (3, 10):
        bad += lib.hostcheck_integrate_v4(1200, seed, mu, voxel, dw, maxw, ident, 1, st)
    voxels, changed, flips, gate = st[0], st[4], st[6], st[7]
    assert bad == 0
    assert flips < 0.02 * changed and gate < 0.002 * changed
