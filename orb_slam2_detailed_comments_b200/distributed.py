"""Multi-GPU plumbing (one process per GPU, torch.distributed): SURVEY.md §8(e).

* Extraction and pairwise matching shard by independent units (frames, frame pairs):
  contiguous index blocks per rank, NO data-path collective.
* All-pairs keyframe matching has one real exchange step: every rank owns the descriptors of
  nKF/world keyframes (its rows) and needs everyone's (the columns). The exchange is an
  all-gather over NCCL/NVLink; `allpairs_match_counts` overlaps it with compute by issuing one
  asynchronous broadcast per peer block and running the all-pairs kernel on each column block
  as soon as it has landed (own block first), on a separate compute stream.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from ._lib import check, lib, ptr, stream_arg


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


class NcclCommunicator:
    """An ncclComm_t made through the C ABI (orb_nccl_unique_id / orb_nccl_comm_create): rank 0 draws the unique id and
    the 128 bytes travel through the existing torch.distributed group (any backend). The library binds NCCL at run time,
    i.e. the copy PyTorch has already loaded. A C++ host passes its own ncclComm_t to orb_match_allpairs_nccl instead."""

    def __init__(self, device, group=None):
        self._L = lib()
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        uid = np.zeros(128, np.uint8)
        if self.rank == 0:
            check(self._L.orb_nccl_unique_id(ptr(uid)))
        backend = dist.get_backend(group)
        t = torch.from_numpy(uid)
        if backend == "nccl":
            t = t.cuda(device)
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast(t, src=src, group=group)
        uid = t.cpu().numpy().copy()
        self.comm = C.c_void_p()
        check(self._L.orb_nccl_comm_create(int(device), self.rank, self.world, ptr(uid), C.byref(self.comm)))

    def close(self):
        if self.comm:
            self._L.orb_nccl_comm_destroy(self.comm)
            self.comm = C.c_void_p()


def allpairs_match_counts_nccl(matcher, comm, local_desc, n_kf, all_desc=None, out=None, stream=None):
    """The all-pairs workload through the library's own multi-GPU entry point (orb_match_allpairs_nccl): per-owner
    ncclBroadcast on the matcher's communication stream, k_allpairs per landed block on the compute stream.
    local_desc: (rows_local, n_desc, 32) u8 CUDA tensor, rows = shard_range(n_kf, rank, world). comm: NcclCommunicator or
    None for a single process. Returns the rank's (rows_local, n_kf) int32 block (asynchronous on `stream`, default
    torch's current stream)."""
    rank = comm.rank if comm is not None else 0
    world = comm.world if comm is not None else 1
    rb, re = shard_range(n_kf, rank, world)
    assert local_desc.shape[0] == re - rb and local_desc.is_cuda and local_desc.is_contiguous()
    n_desc = local_desc.shape[1]
    if all_desc is None:
        all_desc = local_desc.new_empty((n_kf, n_desc, 32))
    if out is None:
        out = torch.empty((re - rb, n_kf), dtype=torch.int32, device=local_desc.device)
    check(matcher._L.orb_match_allpairs_nccl(matcher._h, comm.comm if comm is not None else None, rank, world, ptr(local_desc),
                                             int(n_kf), int(n_desc), matcher.mfNNratio, ptr(all_desc), ptr(out),
                                             matcher._stream(stream, local_desc)))
    return out
