"""Multi-GPU tests (need >= 2 B200s on the box; skipped otherwise): NCCL all-pairs exchange with
the real all-pairs kernel, and frame-sharded extraction, both against the oracle."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from oracle import orb_oracle as O
    from orb_slam2_detailed_comments_b200 import ORBextractor, ORBmatcher
    from orb_slam2_detailed_comments_b200.distributed import (NcclCommunicator, allpairs_match_counts, allpairs_match_counts_nccl,
                                                              shard_range)
    from orb_slam2_detailed_comments_b200.synth import synth_batch

    # ---- all-pairs with the NCCL exchange
    n_kf, nd = 9, 256
    rng = np.random.RandomState(11)
    base = rng.randint(0, 256, (nd, 32)).astype(np.uint8)
    bits = np.unpackbits(base, axis=1)
    all_np = np.stack([np.packbits(bits ^ (rng.rand(*bits.shape) < 0.03 * (k % 3)).astype(np.uint8), axis=1)[rng.permutation(nd)]
                       for k in range(n_kf)])
    rb, re = shard_range(n_kf, rank, world)
    matcher = ORBmatcher(0.9, True, device=rank)
    # kernels must run on the stream torch / NCCL order their work on: a dedicated torch stream
    ts = torch.cuda.Stream(device=dev)
    stream = ts.cuda_stream

    def compute_block(all_desc, r0, r1, c0, c1, out):
        matcher.match_allpairs_device(all_desc, r0, r1, out, stream=stream, col_begin=c0, col_end=c1)

    ok = True
    local = torch.from_numpy(all_np[rb:re].copy()).to(dev)
    torch.cuda.synchronize()
    for overlap in (True, False):
        with torch.cuda.stream(ts):
            counts = allpairs_match_counts(local, n_kf, compute_block, overlap=overlap)
        torch.cuda.synchronize()
        ok = ok and bool(np.array_equal(counts.cpu().numpy(), O.allpairs_counts(all_np, 0.9, rb, re)))

    # ---- the same through the library's own multi-GPU entry point (orb_match_allpairs_nccl: exchange inside the C ABI)
    comm = NcclCommunicator(rank)
    for rep in range(2):
        with torch.cuda.stream(ts):
            counts = allpairs_match_counts_nccl(matcher, comm, local, n_kf)
        torch.cuda.synchronize()
        ok = ok and bool(np.array_equal(counts.cpu().numpy(), O.allpairs_counts(all_np, 0.9, rb, re)))
    comm.close()

    # ---- frame-sharded extraction: each rank extracts its contiguous block of the same 6 frames
    imgs = synth_batch(752, 480, 6, seed0=300)
    fb, fe = shard_range(len(imgs), rank, world)
    ext = ORBextractor(1200, 1.2, 8, 20, 7, device=rank, max_batch=4)
    kps, desc, cnt = ext.extract_batch_host(imgs[fb:fe])
    orc = O.OracleExtractor(1200, 1.2, 8, 20, 7)
    ok_ext = True
    for i in range(fe - fb):
        okps, odesc = orc(imgs[fb + i])
        n = cnt[i]
        ok_ext = ok_ext and n == len(okps) and bool(np.array_equal(kps[i, :n]["x"], okps["x"])) and \
            bool(np.array_equal(kps[i, :n]["y"], okps["y"])) and bool((desc[i, :n] == odesc).all(1).mean() >= 0.999)
    q.put((rank, ok, ok_ext))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_allpairs_and_sharded_extraction():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, ok_ext in res:
        assert ok, "all-pairs block differs from the oracle on rank %d" % rank
        assert ok_ext, "sharded extraction differs from the oracle on rank %d" % rank
