"""Multi-GPU plumbing (one process per GPU, torch.distributed): SURVEY.md §8(e).

* Extraction and pairwise matching shard by independent units (frames, frame pairs):
  contiguous index blocks per rank, NO data-path collective.
* All-pairs keyframe matching has one real exchange step: every rank owns the descriptors of
  nKF/world keyframes (its rows) and needs everyone's (the columns). The exchange is an
  all-gather over NCCL/NVLink; `allpairs_match_counts` overlaps it with compute by issuing one
  asynchronous broadcast per peer block and running the all-pairs kernel on each column block
  as soon as it has landed (own block first), on a separate compute stream.
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous block [begin, end) of `total` units for `rank` (blocks differ by at most one)."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allgather_blocks(local, counts, group=None):
    """All-gather of variable-sized leading-dimension blocks. `counts[r]` = rows owned by rank r.
    Returns the concatenated tensor (sum(counts), ...). Works with nccl (GPU) and gloo (CPU)."""
    world = dist.get_world_size(group)
    if len(set(counts)) == 1:
        out = local.new_empty((sum(counts),) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # uneven blocks: pad every block to the largest, gather, drop the padding
    mx = max(counts)
    padded = local.new_zeros((mx,) + tuple(local.shape[1:]))
    padded[: local.shape[0]].copy_(local)
    out = local.new_empty((world * mx,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * mx: r * mx + counts[r]] for r in range(world)], 0)


def allpairs_match_counts(local_desc, n_kf, compute_block, group=None, overlap=True):
    """local_desc: (rows_local, n_desc, 32) u8 descriptors of this rank's keyframes (block
    `shard_range(n_kf, rank, world)`). compute_block(all_desc, row_begin, row_end, col_begin,
    col_end, out) must fill out[:, col_begin:col_end] for the rank's rows; on the GPU this is
    ORBmatcher.match_allpairs_device. compute_block MUST enqueue its work on torch's CURRENT stream: `out`, `all_desc`
    and the NCCL broadcasts (work.wait() only blocks the current stream) are ordered on it. match_allpairs_device does
    so by default (stream=None -> torch.cuda.current_stream(), see _lib.stream_arg); pass an explicit stream only if it
    is the current one. Returns the rank's (rows_local, n_kf) int32 count block."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    ranges = [shard_range(n_kf, r, world) for r in range(world)]
    rb, re = ranges[rank]
    assert local_desc.shape[0] == re - rb
    out = torch.zeros((re - rb, n_kf), dtype=torch.int32, device=local_desc.device)
    if world == 1:
        compute_block(local_desc, rb, re, 0, n_kf, out)
        return out
    all_desc = local_desc.new_empty((n_kf,) + tuple(local_desc.shape[1:]))
    all_desc[rb:re].copy_(local_desc)
    if not overlap:
        gathered = allgather_blocks(local_desc, [b - a for a, b in ranges], group)
        compute_block(gathered, rb, re, 0, n_kf, out)
        return out
    # one async broadcast per peer block; compute own block first, then blocks in arrival order
    works = []
    for r in range(world):
        a, b = ranges[r]
        src = dist.get_global_rank(group, r) if group is not None else r
        works.append(dist.broadcast(all_desc[a:b], src=src, group=group, async_op=True))
    compute_block(all_desc, rb, re, rb, re, out)
    for k in range(1, world):
        r = (rank + k) % world
        works[r].wait()          # on NCCL: makes the current stream wait for that block only
        a, b = ranges[r]
        compute_block(all_desc, rb, re, a, b, out)
    works[rank].wait()
    return out
