"""N>1 host logic on CPU: two gloo ranks each "render" a volume and the renders are gathered on rank 0
(dynslam_b200.multi: ownership + gather). The z-composite itself is a CUDA kernel (b200_composite_instances, checked
against the oracle in tests/test_gpu_frames.py); here the gathered images are composited by the ORACLE and compared
with a literal Python transcription of CompositeColor (DS/InstRecLib/InstanceReconstructor.cpp:873-905)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dynslam_b200 import multi

H, W = 24, 40


def _render(rank):
    rng = np.random.RandomState(100 + rank)
    depth = rng.uniform(1.0, 20.0, size=(H, W)).astype(np.float32)
    depth[rng.rand(H, W) < (0.2 if rank == 0 else 0.6)] = 0.0
    color = rng.randint(0, 256, size=(H, W, 4)).astype(np.uint8)
    return color, depth


def _reference_composite(colors, depths, boost=0.5):
    tc, td = colors[0].copy(), depths[0].copy()
    for sc, sd in zip(colors[1:], depths[1:]):
        for i in range(H):
            for j in range(W):
                if sd[i, j] != 0 and (td[i, j] == 0 or td[i, j] > sd[i, j]):
                    td[i, j] = sd[i, j]
                    for c in range(3):
                        tc[i, j, c] = np.uint8(min(255.0, sc[i, j, c] * (1.0 + boost)))
    return tc, td


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert multi.local_volumes(8, rank, world) == list(range(rank, 8, world))
    color, depth = _render(rank)
    cols, deps = multi.gather_renders(torch.from_numpy(color), torch.from_numpy(depth), dst=0)
    if rank == 0:
        q.put(([c.numpy() for c in cols], [d.numpy() for d in deps]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_and_composite():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    cols, deps = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    renders = [_render(r) for r in range(2)]
    for r in range(2):                                       # the gather delivered every rank's render, in rank order
        assert np.array_equal(cols[r], renders[r][0]) and np.array_equal(deps[r], renders[r][1])
    import ctypes as C
    from dynslam_b200 import abi
    from tests import hostlib
    got_c, got_d = np.ascontiguousarray(cols[0]).copy(), np.ascontiguousarray(deps[0]).copy()
    layers = (abi.InstanceLayer * 1)()
    s_c, s_d = np.ascontiguousarray(cols[1]), np.ascontiguousarray(deps[1])
    layers[0].d_color, layers[0].d_depth = s_c.ctypes.data, s_d.ctypes.data
    hostlib.oracle().oracle_composite_instances(hostlib.vptr(got_c), hostlib.vptr(got_d), H * W, layers, 1, -1.0, 0.0)
    want_c, want_d = _reference_composite([r[0] for r in renders], [r[1] for r in renders])
    assert np.array_equal(got_d, want_d)
    assert np.array_equal(got_c[..., :3], want_c[..., :3])


def test_ownership_and_no_cpu_fallback():
    assert multi.volume_owner(0, 8) == 0 and multi.volume_owner(9, 8) == 1
    t = torch.tensor([[0.0, 2.0, 3.0, 0.0]])
    try:
        multi.composite_depth(None, t, t.clone())
    except RuntimeError as ex:
        assert "no CPU fallback" in str(ex)
    else:
        raise AssertionError("compositing CPU tensors must fail loudly")
