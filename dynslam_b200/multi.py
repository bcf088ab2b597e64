"""One-volume-per-GPU sharding of independent ITMScene volumes and the only exchange step of the path.

The reference keeps every volume (static map + one per car, DS/InstRecLib/InstanceReconstructor.cpp:363-389) on one GPU,
fuses and renders them one after the other and composites the renders on the CPU (CompositeDepth / CompositeColor /
CompositeInstances, InstanceReconstructor.cpp:851-987). Here volume v lives on rank v % world (rank 0 keeps the static
background), fusion / allocation / decay need no communication, and per frame the colour (RGBA8) and depth (f32) renders of
every volume are handed to rank 0 and z-composited there.

`VolumeExchange` is the host side of that step. Its GPU transport is the C++ exchange of libb200fusion (csrc/comm.cu:
NCCL send/recv on the communicator's own stream, double-buffered, composite kernel behind the receives — the engine's
stream never waits for a peer). `GlooTransport` carries the same protocol over torch.distributed's gloo backend with
numpy compositing by the caller's function, so that the orchestration (ownership, slot discipline, layer order, tints) is
tested on CPU with world_size 2 (tests/test_multi_gloo.py); it is test infrastructure, not a fallback of the GPU path.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import abi

# kMatplotlib2Palette (DS/InstRecLib/InstanceReconstructor.cpp:43-55): instance layer tint = palette[track id % 10]
PALETTE = [(0x1f, 0x77, 0xb4, 255), (0xff, 0x7f, 0x0e, 255), (0x2c, 0xa0, 0x2c, 255), (0xd6, 0x27, 0x28, 255),
           (0x94, 0x67, 0xbd, 255), (0x8c, 0x56, 0x4b, 255), (0xe3, 0x77, 0xc2, 255), (0x71, 0x71, 0x71, 255),
           (0xbc, 0xbd, 0x22, 255), (0x17, 0xbe, 0xcf, 255)]


def volume_owner(volume_index, world_size):
    """Rank that owns volume `volume_index` (0 = static background)."""
    return volume_index % world_size


def local_volumes(num_volumes, rank, world_size):
    return [v for v in range(num_volumes) if volume_owner(v, world_size) == rank]


def layer_tints(world_size):
    """flat int32 tints of ranks 1..world-1 in rank order (what b200_gather_composite_submit expects on rank 0)"""
    flat = []
    for r in range(1, world_size):
        flat.extend(PALETTE[(r - 1) % len(PALETTE)])
    return flat


def broadcast_comm_id(make_id, rank, group=None):
    """rank 0 makes the 128-byte rendez-vous id, everybody gets it (any torch.distributed backend)"""
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


class VolumeExchange:
    """Per-frame hand-over of one volume's renders to rank 0 + composite there, two slots deep.

    submit(k, color, depth, out_color=None, out_depth=None) enqueues frame k's exchange; before frame k+2's renders may
    overwrite the same buffers call release(k + 2) (GPU transport: a device-side wait on the engine's stream; gloo: no-op).
    finish() blocks until everything submitted has been delivered (and composited on rank 0)."""

    def __init__(self, engine, img_size, rank, world, dim_factor=0.10, tint_strength=1.0, transport=None):
        self.e, self.rank, self.world = engine, rank, world
        self.w, self.h = img_size
        self.dim_factor, self.tint_strength = float(dim_factor), float(tint_strength)
        self.tints = layer_tints(world)
        self.transport = transport if transport is not None else NcclCxxTransport(engine, img_size, rank, world)

    def submit(self, frame, color, depth, out_color=None, out_depth=None):
        if self.rank == 0 and (out_color is None or out_depth is None):
            raise ValueError("rank 0 composites: it needs out_color / out_depth")
        self.transport.submit(frame & 1, color, depth, out_color, out_depth, self.tints, self.dim_factor, self.tint_strength)

    def release(self, frame):
        self.transport.release(frame & 1)

    def finish(self):
        self.transport.wait(0)
        self.transport.wait(1)

    def close(self):
        self.transport.close()


class NcclCxxTransport:
    """csrc/comm.cu over the C-ABI (b200_comm_*, b200_gather_composite_*)."""

    def __init__(self, engine, img_size, rank, world):
        if not torch.cuda.is_available():
            raise RuntimeError("dynslam_b200.multi exchanges and composites on the GPU only (libb200fusion); there is no CPU fallback")
        self.e, self.lib = engine, abi.load_library()

        def make_id():
            buf = C.create_string_buffer(128)
            if self.lib.b200_comm_unique_id(buf):
                raise RuntimeError(self.lib.b200_comm_last_error(None).decode())
            return buf.raw

        cid = broadcast_comm_id(make_id, rank) if world > 1 else make_id()
        h = C.c_void_p()
        rc = self.lib.b200_comm_create(engine.scene.device.index or 0, world, rank, cid, img_size[0], img_size[1], C.byref(h))
        if rc:
            raise RuntimeError(self.lib.b200_comm_last_error(None).decode())
        self.h = h

    def _check(self, rc):
        if rc:
            raise RuntimeError(self.lib.b200_comm_last_error(self.h).decode())

    def submit(self, slot, color, depth, out_color, out_depth, tints, dim_factor, tint_strength):
        if tints is not getattr(self, "_tints_src", None):        # marshalled once: the tints of a run do not change
            self._tints_src, self._tints_c = tints, ((C.c_int32 * max(len(tints), 1))(*tints) if tints else None)
        t = self._tints_c
        self._check(self.lib.b200_gather_composite_submit(self.h, self.e.h, color.data_ptr(), depth.data_ptr(),
                                                          out_color.data_ptr() if out_color is not None else None,
                                                          out_depth.data_ptr() if out_depth is not None else None, t,
                                                          dim_factor, tint_strength, slot))

    def release(self, slot):
        self._check(self.lib.b200_gather_composite_release(self.h, self.e.h, slot))

    def wait(self, slot):
        self._check(self.lib.b200_gather_composite_wait(self.h, slot))

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200_comm_destroy(self.h)
            self.h = None


class GlooTransport:
    """The same protocol over torch.distributed (gloo) on host tensors; `composite(out_color, out_depth, layers, tints,
    dim_factor, tint_strength)` is supplied by the caller (the CPU tests pass the oracle's). TEST INFRASTRUCTURE."""

    def __init__(self, rank, world, composite, group=None):
        self.rank, self.world, self.composite, self.group = rank, world, composite, group
        self.busy = [False, False]

    def submit(self, slot, color, depth, out_color, out_depth, tints, dim_factor, tint_strength):
        if self.busy[slot]:
            raise RuntimeError(f"slot {slot} resubmitted before release / wait")
        if self.rank != 0:
            dist.send(color.contiguous(), dst=0, group=self.group)
            dist.send(depth.contiguous(), dst=0, group=self.group)
        else:
            layers = []
            for r in range(1, self.world):
                c, d = torch.empty_like(color), torch.empty_like(depth)
                dist.recv(c, src=r, group=self.group)
                dist.recv(d, src=r, group=self.group)
                layers.append((c, d))
            out_color.copy_(color)
            out_depth.copy_(depth)
            self.composite(out_color, out_depth, layers, tints, dim_factor, tint_strength)
        self.busy[slot] = True

    def release(self, slot):
        self.busy[slot] = False

    def wait(self, slot):
        self.busy[slot] = False

    def close(self):
        pass


def composite_all(engine, colors, depths, tints=None, tint_strength=0.0, dim_factor=-1.0):
    """CompositeInstances — InstanceReconstructor.cpp:932-987 on renders that already sit on one GPU: background = volume 0,
    every further volume z-composited on top in order, one kernel (b200_composite_instances). Returns (color, depth)."""
    from . import engine as E
    for t in list(colors) + list(depths):
        if not t.is_cuda:
            raise RuntimeError("dynslam_b200.multi composites on the GPU only (libb200fusion, frames.cu); there is no CPU fallback")
    color, depth = colors[0].clone(), depths[0].clone()
    layers = [(colors[i], depths[i], tints[i] if tints is not None else (0, 0, 0, 0)) for i in range(1, len(colors))]
    E.InstanceFrames(engine).CompositeInstances(color, depth, layers, dim_factor=dim_factor, tint_strength=tint_strength)
    return color, depth
