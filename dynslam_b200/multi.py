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


def composite_depth(target, source):
    """CompositeDepth — InstanceReconstructor.cpp:851-869 (0 = no measurement)."""
    both = (target != 0) & (source != 0)
    out = torch.where(target == 0, source, target)
    return torch.where(both, torch.minimum(target, source), out)


def composite_color(t_color, t_depth, s_color, s_depth, tint=(0, 0, 0, 0), tint_strength=0.0, color_boost=0.5):
    """CompositeColor — InstanceReconstructor.cpp:873-908: the instance wins a pixel when it has depth
    and the target has none or is farther. Returns (color, depth)."""
    on_top = (s_depth != 0) & ((t_depth == 0) | (t_depth > s_depth))
    strength = 1.0 + color_boost - tint_strength
    tint_t = torch.tensor(tint[:3], dtype=torch.float64, device=s_color.device)
    boosted = torch.clamp(s_color[..., :3].to(torch.float64) * strength + tint_t * tint_strength, max=255.0).to(torch.uint8)
    color = t_color.clone()
    color[..., :3] = torch.where(on_top[..., None], boosted, t_color[..., :3])
    depth = torch.where(on_top, s_depth, t_depth)
    return color, depth


def composite_all(colors, depths, tints=None, tint_strength=0.0):
    """Background = volume 0; every further volume is composited on top in rank order."""
    color, depth = colors[0].clone(), depths[0].clone()
    for i in range(1, len(colors)):
        tint = tints[i] if tints is not None else (0, 0, 0, 0)
        color, depth = composite_color(color, depth, colors[i], depths[i], tint, tint_strength)
    return color, depth
