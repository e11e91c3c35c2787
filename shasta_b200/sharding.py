"""Scene sharding across GPUs.

Every (class, frame pair) of the reference's evaluation is independent (tracks are the previous frame's detections,
det3d/datasets/nuscenes/nuscenes.py:213-246), so N ranks run N independent shards of the hot path with replicated
weights and NO data-path collective. Scenes are dealt round-robin so that a scene's frame pairs stay contiguous on
one rank (the per-scene greedy tracker downstream, tools/nusc_shasta/pub_tracker.py, is sequential in time).
The only collective is the gather of fixed-shape per-rank results: NCCL on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def scenes_for_rank(num_scenes, world_size, rank):
    """Round-robin scene ids of this rank."""
    return list(range(rank, num_scenes, world_size))


def frame_pairs_for_rank(scene_lengths, world_size, rank):
    """[(scene, frame_pair_index)] of this rank, scene-major, in time order inside a scene."""
    out = []
    for s in scenes_for_rank(len(scene_lengths), world_size, rank):
        out.extend((s, f) for f in range(scene_lengths[s]))
    return out


def padded_count(scene_lengths, world_size):
    """Largest per-rank number of frame pairs: every rank pads its result block to this many rows so that the
    gather is a single fixed-shape all_gather_into_tensor."""
    return max(len(frame_pairs_for_rank(scene_lengths, world_size, r)) for r in range(world_size))


def batches(items, batch):
    for i in range(0, len(items), batch):
        yield items[i:i + batch]


def gather_rank_blocks(local_block, group=None):
    """all-gather of identically shaped per-rank blocks -> (world, *local_block.shape)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_block.unsqueeze(0)
    local_block = local_block.contiguous()
    shape = tuple(local_block.shape)
    # concatenated layout (world * rows, ...): accepted by both the NCCL and the gloo backend
    out = torch.empty((world * shape[0],) + shape[1:], dtype=local_block.dtype, device=local_block.device)
    dist.all_gather_into_tensor(out, local_block, group=group)
    return out.view((world,) + shape)


def scatter_to_scene_order(gathered, scene_lengths, world_size):
    """Undo the round-robin deal: gathered (world, padded, ...) -> list over scenes of (len_s, ...) tensors."""
    per_scene = [None] * len(scene_lengths)
    for r in range(world_size):
        row = 0
        for s in scenes_for_rank(len(scene_lengths), world_size, r):
            n = scene_lengths[s]
            per_scene[s] = gathered[r, row:row + n]
            row += n
    return per_scene
