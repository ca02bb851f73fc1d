"""Frame sharding + feature-record exchange of the batched stream mode, device agnostic so that the
host logic is testable with the gloo backend on CPU (tests/test_sharding_gloo.py) and runs over NCCL on the GPUs."""
import torch


def shard_range(total_frames, world, rank):
    """Contiguous chunk of frames owned by `rank`: [r*F/G, (r+1)*F/G)."""
    lo = total_frames * rank // world
    hi = total_frames * (rank + 1) // world
    return lo, hi


def predecessor_source(rank, world):
    """Rank whose LAST frame precedes this shard's first frame in the ring-ordered stream."""
    return (rank - 1) % world


def gather_records(kps, desc, counts, g_kps, g_desc, g_counts, group=None):
    """All-gather the fixed-size per-frame records of every rank (one collective per array)."""
    import torch.distributed as dist
    if dist.get_backend(group) == "gloo":  # gloo has no all_gather_into_tensor for all dtypes: use the list form
        for src, dst in ((kps, g_kps), (desc, g_desc), (counts, g_counts)):
            parts = list(dst.unbind(0))
            dist.all_gather(parts, src.contiguous(), group=group)
    else:
        dist.all_gather_into_tensor(g_kps, kps, group=group)
        dist.all_gather_into_tensor(g_desc, desc, group=group)
        dist.all_gather_into_tensor(g_counts, counts, group=group)


def gather_array(x, g_x, group=None):
    """All-gather one more fixed-size per-frame array (e.g. the stereo depths that travel with a feature record)."""
    import torch.distributed as dist
    if dist.get_backend(group) == "gloo":
        dist.all_gather(list(g_x.unbind(0)), x.contiguous(), group=group)
    else:
        dist.all_gather_into_tensor(g_x, x.contiguous(), group=group)


def take_predecessor(g_kps, g_desc, g_counts, rank, world, out_kps, out_desc, out_count):
    """Copy the record of the frame preceding this shard (last frame of the previous rank) into slot 0."""
    prev = predecessor_source(rank, world)
    F = g_kps.shape[1]
    out_kps.copy_(g_kps[prev, F - 1])
    out_desc.copy_(g_desc[prev, F - 1])
    out_count.copy_(g_counts[prev, F - 1:F])
