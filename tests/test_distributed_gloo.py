"""world_size-2 gloo tests (CPU) of the N>1 host logic: frame sharding and the all-pairs
exchange (all-gather of descriptor blocks, row-block compute, overlap schedule). The compute
callback is the oracle here; on GPUs it is ORBmatcher.match_allpairs_device."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from orb_slam2_detailed_comments_b200.distributed import shard_range
    for total in (0, 1, 7, 8, 8192, 4097):
        for world in (1, 2, 3, 8):
            r = [shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_kf, overlap, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import orb_oracle as O
    from orb_slam2_detailed_comments_b200.distributed import allgather_blocks, allpairs_match_counts, shard_range
    rng = np.random.RandomState(5)
    base = rng.randint(0, 256, (40, 32)).astype(np.uint8)
    bits = np.unpackbits(base, axis=1)
    kfs = []
    for k in range(n_kf):
        flips = (rng.rand(*bits.shape) < 0.02 * (k % 3)).astype(np.uint8)
        kfs.append(np.packbits(bits ^ flips, axis=1)[rng.permutation(40)])
    all_np = np.stack(kfs)
    rb, re = shard_range(n_kf, rank, world)
    local = torch.from_numpy(all_np[rb:re].copy())

    def compute_block(all_desc, r0, r1, c0, c1, out):
        full = O.allpairs_counts(all_desc.numpy(), 0.9, r0, r1)
        out[:, c0:c1] = torch.from_numpy(full[:, c0:c1])

    counts = allpairs_match_counts(local, n_kf, compute_block, overlap=overlap)
    gathered = allgather_blocks(local, [shard_range(n_kf, r, world)[1] - shard_range(n_kf, r, world)[0] for r in range(world)])
    ok_gather = bool(np.array_equal(gathered.numpy(), all_np))
    ref = O.allpairs_counts(all_np, 0.9, rb, re)
    q.put((rank, ok_gather, bool(np.array_equal(counts.numpy(), ref)), int(ref.sum())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_kf,overlap", [(6, True), (7, True), (6, False)])
def test_allpairs_exchange_world2(n_kf, overlap):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + n_kf * 7 + int(overlap)) % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_kf, overlap, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_gather, ok_counts, total in res:
        assert ok_gather, "all-gather assembled the keyframes in the wrong order on rank %d" % rank
        assert ok_counts, "row block differs from the oracle on rank %d" % rank
        assert total > 0


def test_c_abi_shard_range_matches_the_python_schedule():
    """orb_shard_range (the block layout orb_match_allpairs_nccl uses) == distributed.shard_range."""
    import ctypes as C
    from orb_slam2_detailed_comments_b200._lib import lib
    from orb_slam2_detailed_comments_b200.distributed import shard_range
    L = lib()
    for total in (7, 8, 9, 1000, 4096):
        for world in (1, 2, 3, 8):
            for rank in range(world):
                a, b = C.c_int(), C.c_int()
                L.orb_shard_range(total, rank, world, C.byref(a), C.byref(b))
                assert (a.value, b.value) == shard_range(total, rank, world)
