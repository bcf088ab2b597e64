"""GPU: marching-cubes meshing (dynslam_b200/csrc/mesh.cu) against the oracle (oracle/mesh_oracle.c, pinned
to the reference's ITMMeshingEngine_CPU in tests/test_oracle_vs_ref.py). Bit-exact, triangle for triangle, in the CPU engine's order;
the OBJ file written from it is byte-identical to one written from the oracle's triangles."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest
import torch

from dynslam_b200 import abi, engine as E, synth
from tests import hostlib as H
from tests import parity as P

pytestmark = pytest.mark.gpu


def _fused_pair(cfg):
    pair = P.Pair(cfg)
    for i, (depth, rgb, M, proj) in enumerate(P.frames_of(cfg)):
        gv, hv = pair.views(depth, rgb, M, proj)
        pair.reco.AllocateSceneFromDepth(pair.scene, gv, pair.rs)
        pair.reco.IntegrateIntoScene(pair.scene, gv, pair.rs)
        assert pair.L.oracle_allocate_from_depth(pair.host.engine, C.byref(pair.host.scene), C.byref(pair.host.rs), C.byref(hv), 0, 0) == 0
        pair.L.oracle_integrate(pair.host.engine, C.byref(pair.host.scene), C.byref(pair.host.rs), C.byref(hv), 0)
    pair.compare_scene("before meshing")
    return pair


def _oracle_mesh(pair, nmax):
    tri = np.zeros(nmax, dtype=abi.TRIANGLE_DTYPE)
    n = pair.L.oracle_mesh_scene(C.byref(pair.host.scene), H.vptr(tri), nmax)
    return n, tri


def test_mesh_scene_bit_exact_and_obj():
    cfg = P.Cfg(frames=5, raycast=False)
    pair = _fused_pair(cfg)
    mesh = E.Mesh(pair.scene)
    n = E.MeshingEngine(pair.eng).MeshScene(mesh, pair.scene)
    no, tri = _oracle_mesh(pair, mesh.noMaxTriangles)
    assert n == no and n > 30000
    got = mesh.to_host()
    assert got.tobytes() == tri[:n].tobytes()
    with tempfile.TemporaryDirectory() as d:                  # ITMMesh::WriteOBJ's format from both triangle lists
        p1, p2 = os.path.join(d, "gpu.obj"), os.path.join(d, "oracle.obj")
        mesh.WriteOBJ(p1)
        m2 = E.Mesh(pair.scene)
        m2.triangles.copy_(torch.from_numpy(tri.view(np.float32).reshape(-1)).to(pair.scene.device))
        m2.noTotalTriangles = no
        m2.WriteOBJ(p2)
        a, b = open(p1, "rb").read(), open(p2, "rb").read()
        assert a == b and a.count(b"\nf ") == n - (0 if a.startswith(b"f ") else 0) and a.startswith(b"v ")


def test_mesh_scene_with_collisions_and_cap():
    """long excess chains (1024 buckets) and a triangle budget smaller than the mesh: the triangles that fit are the first ones
    of the canonical order and the count keeps running (ITMMeshingEngine_CUDA.cu:137-139)."""
    cfg = P.Cfg(frames=4, numBuckets=0x400, excessSize=0x4000, numBlocks=16384, raycast=False)
    pair = _fused_pair(cfg)
    full = E.Mesh(pair.scene)
    n = E.MeshingEngine(pair.eng).MeshScene(full, pair.scene)
    no, tri = _oracle_mesh(pair, full.noMaxTriangles)
    assert n == no and full.to_host().tobytes() == tri[:n].tobytes()
    small = E.Mesh(pair.scene, noMaxTriangles=5000)
    n2 = E.MeshingEngine(pair.eng).MeshScene(small, pair.scene)
    assert n2 == n                                             # the count reports what was generated
    assert small.triangles.cpu().numpy().view(abi.TRIANGLE_DTYPE)[:4999].tobytes() == tri[:4999].tobytes()


def test_mesh_empty_scene():
    cfg = P.Cfg(frames=0)
    pair = P.Pair(cfg)
    mesh = E.Mesh(pair.scene)
    assert E.MeshingEngine(pair.eng).MeshScene(mesh, pair.scene) == 0
