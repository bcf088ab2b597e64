"""CPU-side checks: the C-ABI library loads and exports every symbol include/b200fusion.h
declares (no CUDA call is made), and the host helpers agree with the oracle."""
import ctypes as C
import os
import re
import subprocess

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


def test_ctypes_mirror_has_the_headers_layout(tmp_path):
    """every POD of include/b200fusion.h (+ the diag header's stats), compiled by gcc: size and every field offset must equal the
    ctypes.Structure the Python mirror passes across the boundary"""
    pairs = [("b200_scene", abi.Scene), ("b200_render_state", abi.RenderState), ("b200_view", abi.View), ("b200_camera", abi.Camera),
             ("b200_engine_config", abi.EngineConfig), ("b200_transfer_buffers", abi.TransferBuffers), ("b200_frame_opts", abi.FrameOpts),
             ("b200_view_calib", abi.ViewCalib), ("b200_mask", abi.Mask), ("b200_silhouette_op", abi.SilhouetteOp),
             ("b200_instance_layer", abi.InstanceLayer), ("b200_eval_params", abi.EvalParams), ("b200_eval_callback", abi.EvalCallback),
             ("b200_eval_stats", abi.EvalStats), ("b200_eval_result", abi.EvalResult), ("b200_eval_summary", abi.EvalSummary),
             ("b200_frame_stats", abi.FrameStats)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200fusion.h"', '#include "b200fusion_diag.h"', 'int main(void) {']
    for cname, cls in pairs:
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  printf("b200_triangle size %zu\\n", sizeof(b200_triangle));', '  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(H.ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = {}
    for ln in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines():
        a, b, c = ln.split()
        got[(a, b)] = int(c)
    for cname, cls in pairs:
        assert got[(cname, "size")] == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
    assert got[("b200_triangle", "size")] == abi.TRIANGLE_DTYPE.itemsize == 72


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
