"""One-volume-per-GPU sharding of independent ITMScene volumes and the only exchange step of the path:
gathering the per-volume raycast images on rank 0 and z-compositing them.

The reference keeps every volume (static map + one per car, DS/InstRecLib/InstanceReconstructor.cpp:363-389)
on one GPU, renders them one after the other and composites on the CPU
(CompositeDepth / CompositeColor, InstanceReconstructor.cpp:851-908). Here volume v lives on rank
v % world (rank 0 keeps the background), fusion / allocation / decay need no communication, and per
displayed frame the colour (RGBA8) and depth (f32) renders are gathered over NCCL (gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def volume_owner(volume_index, world_size):
    """Rank that owns volume `volume_index` (0 = static background)."""
    return volume_index % world_size


def local_volumes(num_volumes, rank, world_size):
    return [v for v in range(num_volumes) if volume_owner(v, world_size) == rank]


def gather_renders(color, depth, dst=0, group=None):
    """Gather this rank's (color uint8 [h,w,4], depth float32 [h,w]) on `dst`.
    Returns (list_of_colors, list_of_depths) on dst, (None, None) elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if rank == dst:
        cols = [torch.empty_like(color) for _ in range(world)]
        deps = [torch.empty_like(depth) for _ in range(world)]
    else:
        cols = deps = None
    dist.gather(color, cols, dst=dst, group=group)
    dist.gather(depth, deps, dst=dst, group=group)
    return cols, deps


def _need_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("dynslam_b200.multi composites on the GPU only (libb200fusion, frames.cu); there is no CPU fallback")


def composite_depth(engine, target, source):
    """CompositeDepth — InstanceReconstructor.cpp:850-869, in place on `target` (0 = no measurement)."""
    from . import engine as E
    _need_cuda(target, source)
    E.InstanceFrames(engine).CompositeDepth(target, source)
    return target


def composite_all(engine, colors, depths, tints=None, tint_strength=0.0, dim_factor=-1.0):
    """CompositeInstances — InstanceReconstructor.cpp:932-987 on the gathered renders: background = volume 0, every
    further volume z-composited on top in rank order, one kernel (b200_composite_instances). Returns (color, depth)."""
    from . import engine as E
    _need_cuda(*colors, *depths)
    color, depth = colors[0].clone(), depths[0].clone()
    layers = [(colors[i], depths[i], tints[i] if tints is not None else (0, 0, 0, 0)) for i in range(1, len(colors))]
    E.InstanceFrames(engine).CompositeInstances(color, depth, layers, dim_factor=dim_factor, tint_strength=tint_strength)
    return color, depth
