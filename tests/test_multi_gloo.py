"""N>1 host logic on CPU (world_size 2, gloo): the per-frame exchange protocol of dynslam_b200.multi.VolumeExchange —
ownership, two-slot discipline over several frames, layer order and tints — with GlooTransport standing in for the C++
NCCL exchange (csrc/comm.cu). Rank 0 composites with the ORACLE (oracle_composite_instances, pinned to the reference's
CompositeColor in tests/test_frames_oracle.py) and the result is compared with a literal Python transcription of
CompositeInstances (DS/InstRecLib/InstanceReconstructor.cpp:873-905, :932-987). The CUDA composite kernel itself is checked
against the oracle in tests/test_gpu_frames.py."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dynslam_b200 import multi

H, W, FRAMES = 24, 40, 5


def _render(rank, frame):
    rng = np.random.RandomState(100 + rank + 17 * frame)
    depth = rng.uniform(1.0, 20.0, size=(H, W)).astype(np.float32)
    depth[rng.rand(H, W) < (0.2 if rank == 0 else 0.6)] = 0.0
    color = rng.randint(0, 256, size=(H, W, 4)).astype(np.uint8)
    return color, depth


def _reference_composite(colors, depths, tints, dim_factor, tint_strength):
    """CompositeInstances as the reference writes it: dim the background, then CompositeColor per instance."""
    tc, td = colors[0].copy(), depths[0].copy()
    tc[..., :3] = (tc[..., :3].astype(np.float64) * (1.0 - np.float64(np.float32(dim_factor)))).astype(np.uint8)
    col_strength = 1.0 + float(np.float32(0.50)) - float(np.float32(tint_strength))
    for (sc, sd), tint in zip(zip(colors[1:], depths[1:]), tints):
        for i in range(H):
            for j in range(W):
                if sd[i, j] != 0 and (td[i, j] == 0 or td[i, j] > sd[i, j]):
                    td[i, j] = sd[i, j]
                    for c in range(3):
                        v = float(sc[i, j, c]) * col_strength + float(np.float32(np.float32(tint[c]) * np.float32(tint_strength)))
                        tc[i, j, c] = np.uint8(v if v < 255.0 else 255.0)
    return tc, td


def _oracle_composite(out_color, out_depth, layers, tints, dim_factor, tint_strength):
    import ctypes as C
    from dynslam_b200 import abi
    from tests import hostlib
    oc, od = out_color.numpy(), out_depth.numpy()
    arr = (abi.InstanceLayer * max(len(layers), 1))()
    keep = []
    for k, (c, d) in enumerate(layers):
        cn, dn = np.ascontiguousarray(c.numpy()), np.ascontiguousarray(d.numpy())
        keep.append((cn, dn))
        arr[k].d_color, arr[k].d_depth = cn.ctypes.data, dn.ctypes.data
        arr[k].tint = (C.c_int32 * 4)(*tints[4 * k:4 * k + 4])
    hostlib.oracle().oracle_composite_instances(hostlib.vptr(oc), hostlib.vptr(od), H * W, arr, len(layers), float(dim_factor), float(tint_strength))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert multi.local_volumes(8, rank, world) == list(range(rank, 8, world))
    cid = multi.broadcast_comm_id(lambda: b"x" * 128, rank)
    assert cid == b"x" * 128
    ex = multi.VolumeExchange(None, (W, H), rank, world, dim_factor=0.10, tint_strength=1.0,
                              transport=multi.GlooTransport(rank, world, _oracle_composite))
    outs = []
    out_c = [torch.zeros((H, W, 4), dtype=torch.uint8) for _ in range(2)]
    out_d = [torch.zeros((H, W), dtype=torch.float32) for _ in range(2)]
    for k in range(FRAMES):
        ex.release(k)                                  # slot k & 1 was handed over with frame k - 2
        color, depth = _render(rank, k)
        ex.submit(k, torch.from_numpy(color), torch.from_numpy(depth), out_c[k & 1] if rank == 0 else None,
                  out_d[k & 1] if rank == 0 else None)
        if rank == 0:
            outs.append((out_c[k & 1].numpy().copy(), out_d[k & 1].numpy().copy()))
    ex.finish()
    try:                                               # slot discipline: a third submit without release must fail loudly
        ex.submit(FRAMES, torch.zeros((H, W, 4), dtype=torch.uint8), torch.zeros((H, W)), out_c[0], out_d[0])
        ex.submit(FRAMES + 2, torch.zeros((H, W, 4), dtype=torch.uint8), torch.zeros((H, W)), out_c[0], out_d[0])
        ok = False
    except RuntimeError:
        ok = True
    if rank == 0:
        q.put((outs, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_protocol_and_composite():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs, slot_guard = q.get(timeout=180)
    for p in procs:
        p.join(timeout=120)
    assert slot_guard
    tints = [multi.PALETTE[0]]
    assert multi.layer_tints(2) == list(multi.PALETTE[0])
    for k in range(FRAMES):
        renders = [_render(r, k) for r in range(2)]
        want_c, want_d = _reference_composite([r[0] for r in renders], [r[1] for r in renders], tints, 0.10, 1.0)
        got_c, got_d = outs[k]
        assert np.array_equal(got_d, want_d), k
        assert np.array_equal(got_c[..., :3], want_c[..., :3]), k


def test_ownership_and_no_cpu_fallback():
    assert multi.volume_owner(0, 8) == 0 and multi.volume_owner(9, 8) == 1
    assert len(multi.layer_tints(8)) == 28
    t = torch.zeros((2, 2, 4), dtype=torch.uint8)
    try:
        multi.composite_all(None, [t, t], [torch.zeros((2, 2)), torch.zeros((2, 2))])
    except RuntimeError as ex:
        assert "no CPU fallback" in str(ex)
    else:
        raise AssertionError("compositing CPU tensors must fail loudly")
    if not torch.cuda.is_available():
        try:
            multi.NcclCxxTransport(None, (4, 4), 0, 1)
        except RuntimeError as ex:
            assert "no CPU fallback" in str(ex)
        else:
            raise AssertionError("the GPU transport must refuse to run without a GPU")
