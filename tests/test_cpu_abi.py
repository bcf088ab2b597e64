"""CPU-side checks: the C-ABI library loads and exports every symbol include/b200fusion.h
declares (no CUDA call is made), and the host helpers agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np

from dynslam_b200 import abi, synth
from tests import hostlib as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = abi.load_library()
    hdr = open(os.path.join(ROOT, "include", "b200fusion.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(abi.EXPORTS), declared ^ set(abi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    # the measurement / test hooks live in their own header, outside the drop-in boundary
    diag = open(os.path.join(ROOT, "include", "b200fusion_diag.h")).read()
    declared_diag = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", diag))
    assert declared_diag == set(abi.DIAG_EXPORTS), declared_diag ^ set(abi.DIAG_EXPORTS)
    for name in declared_diag:
        assert hasattr(lib, name), name
    assert not (declared & declared_diag)


def test_pod_layouts_match_header():
    assert C.sizeof(abi.Scene) == 5 * 8 + 12 * 4
    assert abi.HASH_ENTRY_DTYPE.itemsize == 20 and abi.VOXEL_DTYPE.itemsize == 8
    assert abi.HASH_ENTRY_DTYPE.fields["offset"][1] == 8 and abi.HASH_ENTRY_DTYPE.fields["ptr"][1] == 12
    assert abi.VOXEL_DTYPE.fields["w_depth"][1] == 2 and abi.VOXEL_DTYPE.fields["w_color"][1] == 6


def test_host_matrix_helpers_equal_oracle():
    lib, L = abi.host_library(), H.oracle()
    rng = np.random.default_rng(3)
    mats = [synth.kitti_pose(i * 11) for i in range(20)]
    mats += [(rng.standard_normal((4, 4)) * s).astype(np.float32) for s in (1e-3, 1.0, 50.0) for _ in range(200)]
    for m in mats:
        c = abi.mat_to_c(m)
        a, b = abi.f16(), abi.f16()
        assert lib.b200h_mat4_inv(c, a) == L.oracle_mat4_inv(c, b) == 1   # the oracle's is pinned to ORUtils' inv() (test_oracle_vs_ref)
        assert bytes(a) == bytes(b)
        m1, m2 = abi.f16(), abi.f16()
        lib.b200h_mat4_mul(c, a, m1)
        L.oracle_mat4_mul(c, a, m2)
        assert bytes(m1) == bytes(m2)
    z = abi.f16(*([0.0] * 16))
    assert lib.b200h_mat4_inv(z, abi.f16()) == L.oracle_mat4_inv(z, abi.f16()) == 0


def test_engine_refuses_to_run_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dynslam_b200 import engine as E
    with pytest.raises(RuntimeError):
        E.Engine(object(), (64, 64))
