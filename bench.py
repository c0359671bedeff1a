#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native ORB front-end.

Metric (BASELINE.json): ORB-extract frames/s @1241x376, 2000 features (KITTI stereo shape), plus
Hamming comparisons/s of the SearchForInitialization-style brute-force 2-NN matcher.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one pass of the extractor over one batch of 1024 stereo pairs (2048 eye-frames,
955 MB, larger than L2) per GPU. `value` = eye-frames/s with frames resident in HBM; `e2e` =
the same through the C-ABI host entry point (orb_extract_batch_host) with pinned host buffers,
H2D and D2H inside the timed region. One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# workload shapes of the reference's shipped configs (Examples/*/*.yaml): name -> (W, H, nFeatures, bf, fx)
WORKLOADS = {"kitti": (1241, 376, 2000, 386.1448, 718.856),       # Examples/Stereo/KITTI00-02.yaml
             "euroc": (752, 480, 1200, 47.90639384423901, 435.2046959714599),   # Examples/Stereo/EuRoC.yaml
             "tum1": (640, 480, 1000, 40.0, 517.306408)}           # Examples/Monocular/TUM1.yaml (RGB-D bf)
W, H, NFEAT = 1241, 376, 2000          # the headline workload (BASELINE.json configs[1]); --workload overrides
PAIRS_PER_STEP = 1024                  # stereo pairs per step and GPU (BASELINE.json configs[1])
UNIQUE_FRAMES = 512                    # unique synthetic frames per rank (241 MB > 126 MB L2), tiled
MATCH_PAIRS, MATCH_N = 4096, 2000      # BASELINE.json configs[3]
METRIC = "ORB-extract frames/s @1241x376 2k feats"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout must carry exactly ONE JSON line: keep the real stdout aside and point fd 1 at stderr, so
# that anything a library prints to stdout (e.g. NCCL's version banner) cannot pollute it
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def gen_frames(count, seed0):
    import multiprocessing as mp

    import numpy as np

    from orb_slam2_detailed_comments_b200.synth import synth_frame
    procs = max(1, min(16, (os.cpu_count() or 2) // max(1, int(os.environ.get("WORLD_SIZE", "1")))))
    args = [(W, H, seed0 + i) for i in range(count)]
    if procs > 1:
        with mp.get_context("fork").Pool(procs) as pool:
            frames = pool.starmap(synth_frame, args, chunksize=8)
    else:
        frames = [synth_frame(*a) for a in args]
    return np.stack(frames)


def algorithmic_bytes(levels, mean_kp, mean_cand):
    """bytes per frame of the whole path and per stage."""
    in_b = W * H
    bordered = sum((w + 38) * (h + 38) for w, h in levels)
    area = sum(w * h for w, h in levels)
    total = in_b + bordered + 2 * area + mean_kp * (749 + 512) + mean_kp * 56
    per_stage = {
        "pyramid": in_b + bordered + sum(w * h for w, h in levels[:-1]),
        "fast": area + 8 * mean_cand,
        "quadtree": 8 * mean_cand + 8 * mean_kp,
        "blur": 2 * area,
        "describe": mean_kp * (749 + 512) + mean_kp * 60,
    }
    return total, per_stage


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            txt = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            return out
        sm, reasons, smax = [], set(), None
        for line in txt.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm))
        return out


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def issue_roofline(fps_per_gpu, clocks):
    """The extraction path is bound by instruction issue, not by HBM: warp instructions per frame (committed ncu capture of one
    256-frame chunk, profiles/traffic.json) x frames/s against the SM's issue rate (4 warp instructions per clock and SM)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))
        wi = {k: v for k, v in t["warp_instructions"].items() if not k.startswith("_")}
        per_frame = sum(wi.values()) / float(t.get("_frames_per_launch", 256))
        mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
        peak = 4.0 * 148 * mhz * 1e6
        return {"bound": "issue", "warp_instructions_per_frame": per_frame, "achieved": per_frame * fps_per_gpu, "peak": peak,
                "unit": "warp instructions/s", "frac": per_frame * fps_per_gpu / peak,
                "peak_def": "4 warp instructions per clock per SM x 148 SMs x the SM clock sampled during the timed region",
                "source": "profiles/traffic.json (smsp__inst_executed.sum per stage, ncu --set full)"}
    except Exception:
        return None


def ncu_traffic(stage):
    """dram bytes per launch of a stage's kernel from the committed ncu --set full capture, or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(stage)
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------------------------
def cpu_extractor_arm():
    """(kind, description, fn(frames, nthreads) -> total keypoints) of the CPU arm: the reference's own
    src/ORBextractor.cc (oracle/_ref/liborbref.so, compiled unmodified against the OpenCV stand-in of
    oracle/cvshim/, plain malloc) when that binary is present, else the oracle port."""
    from oracle import orb_oracle as O
    from oracle import orb_ref as R
    if R.available():
        def fn(frames, nthreads):
            return R.extract_batch_mt((NFEAT, 1.2, 8, 20, 7), frames, nthreads)
        return ("reference", "the reference's own ORBextractor.cc compiled unmodified (oracle/_ref); its OpenCV calls are "
                "served by the scalar stand-in of oracle/cvshim (the reference with a SIMD OpenCV build would be faster "
                "in resize / FAST / blur; the control flow, quadtree, orientation and descriptors are the reference's code)", fn)

    def fn(frames, nthreads):
        return O.extract_batch_mt(frames, NFEAT, nthreads=nthreads)[0]
    return "port", "CPU oracle port of the reference path (oracle/_ref not built)", fn


def run_reference(args, rank):
    """The reference's CPU implementation of the path: oracle/_ref (the reference's ORBextractor.cc itself) when
    it was built, else the oracle port; all host threads, one frame per thread, on a bounded sample per step."""
    if rank != 0:
        return
    from oracle import orb_oracle as O
    kind, what, cpu_extract = cpu_extractor_arm()
    cores = O.hardware_threads()
    sample = max(64, 8 * cores)
    frames = gen_frames(min(sample, 256), 0)
    if len(frames) < sample:
        import numpy as np
        frames = np.concatenate([frames] * ((sample + len(frames) - 1) // len(frames)))[:sample]
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_extract(frames[: 2 * cores], cores)
    t0 = time.perf_counter()
    kp = 0
    for _ in range(args.steps):
        kp += cpu_extract(frames, cores)
    dt = time.perf_counter() - t0
    fps = args.steps * sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "%s-shape %dx%d eye-frames, ORBextractor(%d,1.2,8,20,7) on the CPU: %s" % (args.workload, W, H, NFEAT, what),
                   "frames_per_step": sample},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": "%d frames per step, one frame per thread" % sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mean_keypoints": kp / (args.steps * sample),
    }
    emit(line)


# --------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(props):
    """N > 1: run this rank on the CPUs next to its GPU, so that its pinned staging memory (first touch) and the
    copy-engine traffic stay on the GPU's NUMA node instead of crossing the socket link. Best effort."""
    try:
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if node >= 0 and cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    # no NUMA information (containers often hide it): give every rank its own contiguous block of the visible CPUs, in
    # GPU order - CPU numbering is socket-contiguous on the usual two-socket boxes and GPUs 0..N/2-1 hang off socket 0, so
    # pinned staging memory (first touch) and the copy threads of a rank stay on one socket instead of wandering
    try:
        world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("LOCAL_RANK", "0"))
        avail = sorted(os.sched_getaffinity(0))
        if world > 1 and len(avail) >= 2 * world:
            per = len(avail) // world
            os.sched_setaffinity(0, set(avail[rank * per:(rank + 1) * per]))
            return "cpu-block %d-%d" % (avail[rank * per], avail[(rank + 1) * per - 1])
    except Exception:
        pass
    return None


def numa_node_of(t):
    """NUMA node of the first pages of a pinned host tensor (move_pages query), or None where the kernel does not tell."""
    try:
        import ctypes
        libc = ctypes.CDLL(None, use_errno=True)
        n = 8
        page = os.sysconf("SC_PAGESIZE")
        base = t.data_ptr() & ~(page - 1)
        pages = (ctypes.c_void_p * n)(*[base + i * page * 64 for i in range(n)])
        status = (ctypes.c_int * n)()
        if libc.syscall(279, 0, n, pages, None, status, 0) != 0:   # __NR_move_pages (x86-64), nodes = NULL: query only
            return None
        nodes = sorted(set(int(v) for v in status if v >= 0))
        return nodes if nodes else None
    except Exception:  # noqa: BLE001
        return None


def parity_counters(ext, frames):
    """north_star: "the only allowed differences are cvRound flips ... these must be counted and reported". A sample of the
    pool's frames is extracted once more through the drop-in call and compared with the CPU oracle, OUTSIDE any timed
    region: keypoint records (x, y, octave, response, size) must be identical, angles within 1e-3 degrees; descriptor rows
    that differ and the bits flipped in them are counted; so are the (level, frame) quadtree runs whose result depends on
    the tie rule for equal-sized nodes (heap-address order in the reference, creation order here and in the oracle)."""
    import numpy as np

    from oracle import orb_oracle as O
    orc = O.OracleExtractor(NFEAT, 1.2, 8, 20, 7)
    out = {"frames_checked": len(frames), "keypoints_compared": 0, "keypoint_records_differing": 0, "max_angle_err_deg": 0.0,
           "desc_rows_compared": 0, "desc_rows_differing": 0, "desc_bits_flipped": 0, "tie_sensitive_levels": 0, "levels_checked": 0,
           "checked_against": "CPU oracle (oracle/orb_oracle.cpp, pinned bit-exactly to the reference's own code in oracle/_ref)"}
    for img in frames:
        kps, desc = ext(img)
        okps, odesc = orc(img)
        n = min(len(kps), len(okps))
        out["keypoints_compared"] += max(len(kps), len(okps))
        bad = np.zeros(n, bool)
        for f in ("x", "y", "octave", "response", "size"):
            bad |= kps[f][:n] != okps[f][:n]
        out["keypoint_records_differing"] += int(bad.sum()) + abs(len(kps) - len(okps))
        if n:
            d = np.abs(kps["angle"][:n] - okps["angle"][:n])
            out["max_angle_err_deg"] = max(out["max_angle_err_deg"], float(np.minimum(d, 360.0 - d).max()))
            x = desc[:n] ^ odesc[:n]
            out["desc_rows_compared"] += n
            out["desc_rows_differing"] += int((x != 0).any(1).sum())
            out["desc_bits_flipped"] += int(np.unpackbits(x).sum())
        for l in range(8):
            out["tie_sensitive_levels"] += int(orc.stats(l)["tie_sensitive"] != 0)
            out["levels_checked"] += 1
    out["desc_identical_fraction"] = 1.0 - out["desc_rows_differing"] / max(1, out["desc_rows_compared"])
    return out


_CV2_FRAMES = None


def _cv2_orb_worker(rng):
    import cv2
    cv2.setNumThreads(1)
    orb = cv2.ORB_create(nfeatures=NFEAT, scaleFactor=1.2, nlevels=8, edgeThreshold=19, fastThreshold=20, scoreType=cv2.ORB_FAST_SCORE)
    n = 0
    for i in range(*rng):
        k, _ = orb.detectAndCompute(_CV2_FRAMES[i], None)
        n += len(k)
    return n


def opencv_orb_baseline(frames, cores):
    """A second CPU figure from the REAL (SIMD) OpenCV of this image: cv2.ORB_create(...).detectAndCompute, one process per
    core. It is OpenCV's own ORB (global FAST + retainBest instead of the per-cell threshold fallback and the quadtree), so
    its keypoints are not the reference's; it brackets from below what a well-optimised CPU ORB front-end costs, where the
    reference arm (scalar OpenCV stand-in) brackets from above."""
    global _CV2_FRAMES
    try:
        import multiprocessing as mp

        import cv2
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)}
    _CV2_FRAMES = frames
    per = (len(frames) + cores - 1) // cores
    ranges = [(i, min(i + per, len(frames))) for i in range(0, len(frames), per)]
    with mp.get_context("fork").Pool(len(ranges)) as pool:
        pool.map(_cv2_orb_worker, [(0, 1)] * len(ranges))   # start the workers, load cv2
        t0 = time.perf_counter()
        kp = sum(pool.map(_cv2_orb_worker, ranges))
        dt = time.perf_counter() - t0
    return {"value": len(frames) / dt, "unit": "frames/s", "cores": len(ranges), "mean_keypoints": kp / len(frames),
            "what": "cv2.ORB_create(%d, 1.2, 8, fastThreshold=20, FAST_SCORE).detectAndCompute, OpenCV %s (SIMD build), one process "
                    "per core: OpenCV's own ORB, not ORB-SLAM2's extractor - a lower bracket for the GPU / CPU ratio" % (NFEAT, cv2.__version__)}


def run_sequence_config3(args, rank, world, local_rank, dev, barrier, max_over_ranks):
    """BASELINE.json configs[2]: an EuRoC-shape 752x480 sequence of 8192 frames, 1200 features, STRONG-sharded: the 8192
    frames are split across the ranks (8192 / N each, contiguous blocks, no collective), every rank's share resident in HBM."""
    import numpy as np
    import torch

    from orb_slam2_detailed_comments_b200 import ORBextractor
    from orb_slam2_detailed_comments_b200.distributed import shard_range
    from orb_slam2_detailed_comments_b200.synth import synth_frame
    w, h, nfeat = WORKLOADS["euroc"][:3]
    total = args.sequence_frames
    fb, fe = shard_range(total, rank, world)
    mine = fe - fb
    uniq = min(128, mine)
    import multiprocessing as mp
    procs = max(1, min(16, (os.cpu_count() or 2) // max(1, world)))
    with mp.get_context("fork").Pool(procs) as pool_:
        frames = np.stack(pool_.starmap(synth_frame, [(w, h, 500000 + 1000 * rank + i) for i in range(uniq)], chunksize=8))
    ext = ORBextractor(nfeat, 1.2, 8, 20, 7, device=local_rank, max_batch=args.chunk)
    cap = ext.max_keypoints_for(w, h)
    d_pool = torch.from_numpy(frames).to(dev)
    d_imgs = d_pool.repeat((mine + uniq - 1) // uniq, 1, 1)[:mine].contiguous()
    d_kps = torch.zeros((mine, cap, 28), dtype=torch.uint8, device=dev)
    d_desc = torch.zeros((mine, cap, 32), dtype=torch.uint8, device=dev)
    d_counts = torch.zeros(mine, dtype=torch.int32, device=dev)
    ts = torch.cuda.Stream(device=dev)
    ext.extract_batch_device(d_imgs, d_kps, d_desc, d_counts, stream=ts.cuda_stream)
    ext.synchronize(ts.cuda_stream)
    barrier()
    reps = 3
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    for _ in range(reps):
        ext.extract_batch_device(d_imgs, d_kps, d_desc, d_counts, stream=ts.cuda_stream)
    e1.record(ts)
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1)) / reps
    ext.synchronize(ts.cuda_stream)
    launches = ext.last_launch_count() * (reps + 1)
    mean_kp = float(d_counts.float().mean().item())
    ext.close()
    return {"what": "BASELINE.json configs[2]: EuRoC-shape %dx%d sequence of %d frames, ORBextractor(%d,1.2,8,20,7), frames sharded "
                    "across the GPUs in contiguous blocks, no collective; frames resident in HBM" % (w, h, total, nfeat),
            "frames": total, "frames_per_gpu": mine, "scaling": "strong", "ms_per_sequence": ms,
            "value": total / (ms * 1e-3), "unit": "frames/s", "mean_keypoints": mean_kp, "unique_frames_per_gpu": uniq,
            "l2": "inputs larger than L2 (%d MB per GPU)" % (mine * w * h // 1000000)}, launches



def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = None
    if world > 1:
        numa = bind_to_gpu_numa_node(torch.cuda.get_device_properties(local_rank))
        dist.init_process_group("nccl", device_id=dev)

    from orb_slam2_detailed_comments_b200 import KP_DTYPE, ORBextractor, ORBmatcher, int_pipe_peak

    frames_per_step = 2 * args.pairs
    t_gen = time.perf_counter()
    pool = gen_frames(UNIQUE_FRAMES, 100000 * rank)
    log("[rank %d] generated %d unique frames in %.1fs" % (rank, len(pool), time.perf_counter() - t_gen))

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ext = ORBextractor(NFEAT, 1.2, 8, 20, 7, device=local_rank, max_batch=args.chunk)
    cap = ext.max_keypoints
    d_pool = torch.from_numpy(pool).to(dev)
    d_imgs = d_pool.repeat((frames_per_step + len(pool) - 1) // len(pool), 1, 1)[:frames_per_step].contiguous()
    d_kps = torch.zeros((frames_per_step, cap, 28), dtype=torch.uint8, device=dev)
    d_desc = torch.zeros((frames_per_step, cap, 32), dtype=torch.uint8, device=dev)
    d_counts = torch.zeros(frames_per_step, dtype=torch.int32, device=dev)
    # a dedicated non-default stream: the library treats a NULL stream as "the handle's own",
    # and torch events only see the stream they are recorded on
    tstream = torch.cuda.Stream(device=dev)
    stream = tstream.cuda_stream
    assert stream != 0
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_device():
        ext.extract_batch_device(d_imgs, d_kps, d_desc, d_counts, stream=stream)

    # ---- device-resident timed region (headline `value`) -----------------------------------------
    for _ in range(args.warmup):
        step_device()
    ext.synchronize(stream)
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(tstream)
    for _ in range(args.steps):
        step_device()
    e1.record(tstream)
    torch.cuda.synchronize()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    barrier()
    # ---- the same K steps once more with the chunks serialised on one stream and CUDA events
    # between the stages: each kernel's duration is then its own (used for the roofline)
    ext.set_lanes(1)
    step_device()
    ext.synchronize(stream)
    ext.set_profiling(True)
    ext.stage_times()
    barrier()
    p0 = torch.cuda.Event(enable_timing=True); p1 = torch.cuda.Event(enable_timing=True)
    p0.record(tstream)
    for _ in range(args.steps):
        step_device()
    p1.record(tstream)
    torch.cuda.synchronize()
    ms_serial = max_over_ranks(p0.elapsed_time(p1))
    stages = ext.stage_times()
    ext.set_profiling(False)
    ext.synchronize(stream)
    ext.set_lanes(1 if os.environ.get("ORB_B200_LANES", "2") == "1" else 2)
    step_device()
    ext.synchronize(stream)
    launches_per_step = ext.last_launch_count()
    counts = d_counts.cpu().numpy()
    mean_kp = float(counts.mean())
    fps = world * frames_per_step * args.steps / (ms_total * 1e-3)

    # ---- stereo front-end: both eyes + Frame::ComputeStereoMatches, device resident ---------------
    d_ur = torch.zeros((frames_per_step // 2, cap), dtype=torch.float32, device=dev)
    d_dp = torch.zeros((frames_per_step // 2, cap), dtype=torch.float32, device=dev)
    # a rectified pair with real disparity: right eye = left eye shifted by 16 px (synthetic)
    d_st = d_imgs.clone()
    d_st[1::2, :, :-16] = d_imgs[0::2, :, 16:]
    d_st[1::2, :, -16:] = d_imgs[0::2, :, -16:]
    kitti_bf, kitti_fx = WORKLOADS[args.workload][3], WORKLOADS[args.workload][4]
    s_steps = max(1, min(args.steps, args.e2e_steps))
    ext.extract_stereo_batch_device(d_st, d_kps, d_desc, d_counts, d_ur, d_dp, kitti_bf, kitti_bf / kitti_fx, stream=stream)
    ext.synchronize(stream)
    barrier()
    s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
    s0.record(tstream)
    for _ in range(s_steps):
        ext.extract_stereo_batch_device(d_st, d_kps, d_desc, d_counts, d_ur, d_dp, kitti_bf, kitti_bf / kitti_fx, stream=stream)
    s1.record(tstream)
    torch.cuda.synchronize()
    st_ms = max_over_ranks(s0.elapsed_time(s1))
    ext.synchronize(stream)
    stereo_launches = ext.last_launch_count() * s_steps
    stereo = {"value": world * (frames_per_step // 2) * s_steps / (st_ms * 1e-3), "unit": "stereo pairs/s",
              "ms_per_step": st_ms / s_steps, "steps": s_steps,
              "mean_stereo_points": float((d_ur >= 0).float().sum(dim=1).mean().item()),
              "what": "orb_extract_stereo_batch_device: extraction of both eyes + ComputeStereoMatches (Frame.cc:831)"}
    del d_st

    # ---- Frame helpers after extraction: UndistortKeyPoints + AssignFeaturesToGrid,
    # then one GetFeaturesInArea query per keypoint (window 15 px * scale of its level, as the
    # projection matchers ask), all on the keypoints the stereo step left in HBM
    from orb_slam2_detailed_comments_b200 import frame as F
    from orb_slam2_detailed_comments_b200._lib import AREA_QUERY_DTYPE
    cam = F.camera(kitti_fx, kitti_fx, W / 2.0 - 0.5, H / 2.0 - 0.5, -0.2834, 0.0739, 0.00019, 1.76e-05, 0.0)
    bounds = F.ComputeImageBounds(cam, W, H)
    d_un = torch.zeros_like(d_kps)
    d_cs = torch.zeros((frames_per_step, 64 * 48 + 1), dtype=torch.int32, device=dev)
    d_ci = torch.zeros((frames_per_step, cap), dtype=torch.int32, device=dev)

    def frame_step():
        F.UndistortKeyPoints(d_kps, d_counts, cam, d_un, device=local_rank, stream=stream)
        F.AssignFeaturesToGrid(d_un, d_counts, bounds, d_cs, d_ci, device=local_rank, stream=stream)

    frame_step()
    hk = d_un.cpu().numpy().view(KP_DTYPE).reshape(frames_per_step, cap)
    hc = d_counts.cpu().numpy()
    qf = np.repeat(np.arange(frames_per_step, dtype=np.int32), hc)
    qsel = np.concatenate([hk[f, : hc[f]] for f in range(frames_per_step)])
    sf = ext.GetScaleFactors()
    q = F.make_queries(qf, qsel["x"], qsel["y"], 15.0 * sf[qsel["octave"]], np.maximum(qsel["octave"] - 1, 0), qsel["octave"] + 1)
    d_q = torch.from_numpy(q.view(np.uint8).reshape(len(q), AREA_QUERY_DTYPE.itemsize)).to(dev)
    area_cap = 256
    d_ao = torch.zeros((len(q), area_cap), dtype=torch.int32, device=dev)
    d_ac = torch.zeros(len(q), dtype=torch.int32, device=dev)

    def area_step():
        F.GetFeaturesInArea(d_un, bounds, d_cs, d_ci, d_q, d_ao, d_ac, device=local_rank, stream=stream)

    area_step()
    torch.cuda.synchronize()
    fh = {}
    for name, fn in (("undistort_grid", frame_step), ("features_in_area", area_step)):
        barrier()
        f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
        f0.record(tstream)
        for _ in range(s_steps):
            fn()
        f1.record(tstream)
        torch.cuda.synchronize()
        fh[name] = max_over_ranks(f0.elapsed_time(f1)) / s_steps
    n_kp_total = int(hc.sum())
    frame_helpers = {
        "what": "UndistortKeyPoints + AssignFeaturesToGrid (Frame.cc:724, :399), then one GetFeaturesInArea (Frame.cc:590) per keypoint, device resident",
        "camera": "KITTI focal length with EuRoC-like distortion (k1=-0.2834 k2=0.0739 p1=1.9e-4 p2=1.76e-5) so that the 5-iteration undistortion runs",
        "undistort_grid": {"ms_per_step": fh["undistort_grid"], "frames_per_s": world * frames_per_step / (fh["undistort_grid"] * 1e-3),
                           "keypoints_per_s": world * n_kp_total / (fh["undistort_grid"] * 1e-3),
                           "hbm_GBps": n_kp_total * (28 + 28 + 28 + 8) / (fh["undistort_grid"] * 1e-3) / 1e9},
        "features_in_area": {"ms_per_step": fh["features_in_area"], "queries_per_s": world * len(q) / (fh["features_in_area"] * 1e-3),
                             "mean_results_per_query": float(d_ac.float().mean().item()), "overflowed_queries": int((d_ac > area_cap).sum().item())},
    }
    frame_launches = 3 * s_steps + 3
    del d_q, d_ao, d_ac, d_un, d_cs, d_ci

    # ---- input stage: cvtColor RGB -> gray and remap rectification of the frames of one chunk, plus
    # ComputeDistinctiveDescriptors over 100 k map points; HBM-bound byte kernels, reported against the measured peak
    from orb_slam2_detailed_comments_b200 import input as IN
    in_B = min(256, frames_per_step)
    d_rgb = torch.randint(0, 256, (in_B, H, W, 3), dtype=torch.uint8, device=dev)
    d_g = torch.zeros((in_B, H, W), dtype=torch.uint8, device=dev)
    yy_, xx_ = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=dev), torch.arange(W, dtype=torch.float32, device=dev), indexing="ij")
    rr_ = ((xx_ - W / 2) ** 2 + (yy_ - H / 2) ** 2) / float(W * W)
    d_mx = (xx_ + (xx_ - W / 2) * 0.08 * rr_ + 1.3).contiguous(); d_my = (yy_ + (yy_ - H / 2) * 0.08 * rr_ - 0.7).contiguous()
    d_rect = torch.zeros((in_B, H, W), dtype=torch.uint8, device=dev)
    mp_counts = np.random.RandomState(5).randint(2, 25, 100000).astype(np.int32)
    mp_off = torch.from_numpy(np.concatenate([[0], np.cumsum(mp_counts)]).astype(np.int32)).to(dev)
    mp_desc = torch.randint(0, 256, (int(mp_counts.sum()), 32), dtype=torch.uint8, device=dev)
    mp_best = torch.zeros(len(mp_counts), dtype=torch.int32, device=dev); mp_bd = torch.zeros((len(mp_counts), 32), dtype=torch.uint8, device=dev)
    input_stage = {}
    hbm_peak = measured_hbm_peak()[0]
    for name, fn, nbytes in (("cvt_rgb2gray", lambda: IN.cvtColorGray(d_rgb, IN.RGB2GRAY, d_g, device=local_rank, stream=stream), in_B * H * W * 4),
                             ("remap_linear", lambda: IN.remap(d_imgs[:in_B], d_mx, d_my, d_rect, device=local_rank, stream=stream), in_B * H * W * 2 + H * W * 8),
                             ("distinctive_descriptors", lambda: IN.ComputeDistinctiveDescriptors(mp_desc, mp_off, 32, mp_best, mp_bd, device=local_rank, stream=stream),
                              int(mp_counts.sum()) * 32 + len(mp_counts) * 40)):
        for _ in range(3):
            fn()
        barrier()
        i0 = torch.cuda.Event(enable_timing=True); i1 = torch.cuda.Event(enable_timing=True)
        i0.record(tstream)
        for _ in range(10):
            fn()
        i1.record(tstream)
        torch.cuda.synchronize()
        ms = max_over_ranks(i0.elapsed_time(i1)) / 10
        input_stage[name] = {"ms": ms, "algorithmic_GBps": nbytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / hbm_peak}
    input_stage["cvt_rgb2gray"]["frames_per_s"] = world * in_B / (input_stage["cvt_rgb2gray"]["ms"] * 1e-3)
    input_stage["remap_linear"]["frames_per_s"] = world * in_B / (input_stage["remap_linear"]["ms"] * 1e-3)
    input_stage["distinctive_descriptors"]["map_points_per_s"] = world * len(mp_counts) / (input_stage["distinctive_descriptors"]["ms"] * 1e-3)
    input_stage["what"] = "%d %dx%d frames per call; 100 k map points with 2..24 observations" % (in_B, W, H)
    input_launches = 3 * 13
    del d_rgb, d_g, d_rect, mp_desc

    # ---- tracking matchers: SearchByProjection(CurrentFrame, LastFrame) on 2000 x 2000 frames,
    # projection + grid + search, device resident; batch throughput and single-frame latency
    from orb_slam2_detailed_comments_b200 import search as SR
    from orb_slam2_detailed_comments_b200.synth import tracking_scene
    tr_B, tr_n = 256, 2000
    tr_uniq = [tracking_scene(tr_n, tr_n, 4242 + 17 * rank + i, w=W, h=H, distinct=0.97) for i in range(8)]
    sfs = ext.GetScaleFactors()

    def tile(key, kp=False):
        a = np.stack([sc_[key].view(np.uint8).reshape(tr_n, 28) if kp else np.ascontiguousarray(sc_[key]) for sc_ in tr_uniq])
        return torch.from_numpy(a).to(dev).repeat((tr_B // len(tr_uniq),) + (1,) * (a.ndim - 1)).contiguous()

    t_kps = tile("cur", True); t_desc = tile("cur_desc"); t_ur = tile("uright"); t_occ = tile("occupied0")
    t_last = tile("last", True); t_Xw = tile("Xw"); t_fl = tile("mp_flags"); t_mpd = tile("mp_desc")
    t_T = torch.from_numpy(np.stack([sc_["Tcw"] for sc_ in tr_uniq])).to(dev).repeat(tr_B // len(tr_uniq), 1, 1).contiguous()
    t_cnt = torch.full((tr_B,), tr_n, dtype=torch.int32, device=dev)
    t_dir = torch.zeros(tr_B, dtype=torch.int32, device=dev)
    t_cs = torch.zeros((tr_B, 64 * 48 + 1), dtype=torch.int32, device=dev); t_ci = torch.zeros((tr_B, tr_n), dtype=torch.int32, device=dev)
    t_q = torch.zeros((tr_B, tr_n, 32), dtype=torch.uint8, device=dev)
    t_mk = torch.zeros((tr_B, tr_n), dtype=torch.int32, device=dev); t_mq = torch.zeros((tr_B, tr_n), dtype=torch.int32, device=dev)
    t_nm = torch.zeros(tr_B, dtype=torch.int32, device=dev)
    t_scr = torch.zeros(SR.scratch_bytes(tr_B, tr_n, tr_n), dtype=torch.uint8, device=dev)
    tb = tr_uniq[0]["bounds"]

    def make_track_step(nb):
        # arguments are prepared once: the timed loop only issues the four C-ABI calls
        kps_, cnt_, cs_, ci_, Xw_, fl_, last_, T_, dir_, q_ = (t_kps[:nb], t_cnt[:nb], t_cs[:nb], t_ci[:nb], t_Xw[:nb], t_fl[:nb],
                                                                  t_last[:nb], t_T[:nb], t_dir[:nb], t_q[:nb])
        mpd_, mk_, mq_, nm_ = t_mpd[:nb], t_mk[:nb], t_mq[:nb], t_nm[:nb]
        fr = SR.device_frames(kps_, t_desc[:nb], cnt_, tb, cs_, ci_, t_ur[:nb], t_occ[:nb])

        def step():
            F.AssignFeaturesToGrid(kps_, cnt_, tb, cs_, ci_, device=local_rank, stream=stream)
            SR.ProjectLastFrame(Xw_, fl_, last_, cnt_, T_, dir_, tr_uniq[0]["cam4"], tb, tr_uniq[0]["mbf"], 15.0, sfs, q_,
                                device=local_rank, stream=stream)
            SR.SearchByProjection(fr, q_, mpd_, cnt_, SR.ORB_SEARCH_BEST, SR.TH_HIGH, 0.9, True, t_scr, mk_, mq_, nm_,
                                  device=local_rank, stream=stream)
        return step

    tracking = {"what": "AssignFeaturesToGrid + projection of 2000 last-frame map points + SearchByProjection(CurrentFrame, LastFrame, th=15) "
                        "(ORBmatcher.cc:1710) per frame, device resident; 4 launches per call",
                "scene": "synthetic: 2000 keypoints, 2000 map points of which 97 % follow a keypoint of their own (3 px noise), 8 unique scenes tiled"}
    for name, nb, reps in (("batch", tr_B, 5), ("single_frame", 1, 50)):
        track_step = make_track_step(nb)
        for _ in range(3):
            track_step()
        barrier()
        k0 = torch.cuda.Event(enable_timing=True); k1 = torch.cuda.Event(enable_timing=True)
        k0.record(tstream)
        for _ in range(reps):
            track_step()
        k1.record(tstream)
        torch.cuda.synchronize()
        ms = max_over_ranks(k0.elapsed_time(k1)) / reps
        tracking[name] = {"frames": nb, "ms": ms, "frames_per_s": world * nb / (ms * 1e-3)}
    tracking["mean_matches"] = float(t_nm.float().mean().item())
    track_launches = 4 * (5 + 50 + 6)
    del t_kps, t_desc, t_last, t_Xw, t_mpd, t_q, t_scr

    # ---- drop-in latency: one frame / one stereo pair per call through the reference-shaped entry
    # points (what Frame::Frame does: H2D, all kernels, D2H, synchronise), host wall clock
    lat = {}
    one = pool[0]; two = pool[1]
    for name, fn in () if args.no_latency else (("orb_extract_ms", lambda: ext(one)),
                     ("orb_extract_with_pyramid_ms", lambda: ext(one, want_pyramid=True)),
                     ("orb_extract_stereo_ms", lambda: ext.extract_stereo(one, two, kitti_bf, kitti_bf / kitti_fx))):
        for _ in range(5):
            fn()
        ts_ = []
        for _ in range(30):
            t0 = time.perf_counter(); fn(); ts_.append(time.perf_counter() - t0)
        ts_.sort()
        lat[name] = {"median": 1e3 * ts_[len(ts_) // 2], "p90": 1e3 * ts_[int(len(ts_) * 0.9)]}
    if not args.no_latency:
        # the same call timed INSIDE the C ABI (orb_last_call_breakdown): what a C++ host pays, without the Python wrapper
        acc = {}
        for _ in range(50):
            ext(one)
            for k_, v_ in ext.last_call_breakdown().items():
                acc[k_] = acc.get(k_, 0.0) + v_ / 50
        acc["total_us"] = sum(acc.values())
        lat["orb_extract_inside_c_abi_us"] = acc

    # ---- end to end through the host entry point of the C ABI (pinned buffers) -----------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h_imgs = torch.empty((frames_per_step, H, W), dtype=torch.uint8, pin_memory=True)
    h_imgs.copy_(d_imgs)
    h_kps = torch.empty((frames_per_step, cap, 28), dtype=torch.uint8, pin_memory=True)
    h_desc = torch.empty((frames_per_step, cap, 32), dtype=torch.uint8, pin_memory=True)
    h_counts = torch.empty(frames_per_step, dtype=torch.int32, pin_memory=True)
    np_imgs = h_imgs.numpy(); np_kps = h_kps.numpy().view(KP_DTYPE).reshape(frames_per_step, cap)
    np_desc = h_desc.numpy(); np_counts = h_counts.numpy()
    # a second set of pinned output buffers: steps are issued with orb_extract_batch_host_async, so that the upload
    # of step k+1 overlaps the kernels / download of step k (one continuous pipeline; every step still uploads its
    # frames and downloads its keypoints inside the timed region, and the loop ends with orb_synchronize)
    h_kps2 = torch.empty((frames_per_step, cap, 28), dtype=torch.uint8, pin_memory=True)
    h_desc2 = torch.empty((frames_per_step, cap, 32), dtype=torch.uint8, pin_memory=True)
    h_counts2 = torch.empty(frames_per_step, dtype=torch.int32, pin_memory=True)
    outs = [(np_kps, np_desc, np_counts),
            (h_kps2.numpy().view(KP_DTYPE).reshape(frames_per_step, cap), h_desc2.numpy(), h_counts2.numpy())]
    ext.extract_batch_host_into(np_imgs, np_kps, np_desc, np_counts)  # warm-up (allocates staging)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ext.extract_batch_host_into(np_imgs, np_kps, np_desc, np_counts)
    torch.cuda.synchronize()
    e2e_sync_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        ext.extract_batch_host_into(np_imgs, *outs[i & 1], wait=False)
    ext.synchronize()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_fps = world * frames_per_step * e2e_steps / e2e_s
    e2e_sync_fps = world * frames_per_step * e2e_steps / e2e_sync_s
    e2e_launches = ext.last_launch_count() * e2e_steps * 2
    # parity counters on a sample of the pool, outside every timed region (rank 0)
    parity = parity_counters(ext, pool[: args.parity_frames]) if rank == 0 and args.parity_frames > 0 else None
    assert int(np_counts.sum()) == int(counts.sum()), "host path and device path disagree"
    if e2e_steps > 1:
        assert int(outs[1][2].sum()) == int(counts.sum()), "asynchronous host path and device path disagree"
    del h_kps2, h_desc2
    h2d = frames_per_step * W * H
    # the copy engines alone on the same pinned buffers, all ranks at once (barrier): the ceiling of any end-to-end number
    # on this box. First the upload alone, then upload and download TOGETHER in the byte ratio of a step (two streams), which
    # is the traffic the pipeline really generates (H2D of the frames while the previous chunk's keypoints come back).
    d_probe = torch.empty_like(d_imgs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        d_probe.copy_(h_imgs, non_blocking=True)
    torch.cuda.synchronize()
    h2d_peak = 3 * h2d / max_over_ranks(time.perf_counter() - t0) / 1e9
    s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        with torch.cuda.stream(s_up):
            d_probe.copy_(h_imgs, non_blocking=True)
        with torch.cuda.stream(s_dn):
            h_kps.copy_(d_kps, non_blocking=True)
            h_desc.copy_(d_desc, non_blocking=True)
    torch.cuda.synchronize()
    bidir_s = max_over_ranks(time.perf_counter() - t0)
    bidir_fps = world * 3 * frames_per_step / bidir_s
    pinned_node = numa_node_of(h_imgs)
    del d_probe
    d2h = frames_per_step * (cap * 60 + 4)
    del h_imgs, h_kps, h_desc

    # ---- matching: SearchForInitialization-style brute force 2-NN, 2000 x 2000 per pair ------
    from orb_slam2_detailed_comments_b200.synth import correlated_descriptor_pair
    uniq = 64
    dsc = np.zeros((2 * uniq, MATCH_N, 32), np.uint8); ang = np.zeros((2 * uniq, MATCH_N), np.float32)
    for p in range(uniq):
        A, B, aa, ab = correlated_descriptor_pair(MATCH_N, 7000 + 1000 * rank + p)
        dsc[2 * p], dsc[2 * p + 1], ang[2 * p], ang[2 * p + 1] = A, B, aa, ab
    match_pairs = args.match_pairs
    reps = max(1, match_pairs // uniq)
    match_pairs = reps * uniq
    d_dsc = torch.from_numpy(dsc).to(dev).repeat(reps, 1, 1).contiguous()
    d_ang = torch.from_numpy(ang).to(dev).repeat(reps, 1).contiguous()
    d_m12 = torch.zeros((match_pairs, MATCH_N), dtype=torch.int32, device=dev)
    d_nm = torch.zeros(match_pairs, dtype=torch.int32, device=dev)
    matcher = ORBmatcher(0.9, True, device=local_rank, max_keypoints=MATCH_N)
    m_steps = max(1, min(args.steps, args.match_steps))
    for _ in range(min(args.warmup, 2)):
        matcher.match_pairs_device(d_dsc, d_ang, d_m12, d_nm, stream=stream)
    barrier()
    m0 = torch.cuda.Event(enable_timing=True); m1 = torch.cuda.Event(enable_timing=True)
    m0.record(tstream)
    for _ in range(m_steps):
        matcher.match_pairs_device(d_dsc, d_ang, d_m12, d_nm, stream=stream)
    m1.record(tstream)
    torch.cuda.synchronize()
    m_ms = max_over_ranks(m0.elapsed_time(m1))
    cmp_per_s = world * match_pairs * MATCH_N * MATCH_N * m_steps / (m_ms * 1e-3)
    mean_matches = float(d_nm.float().mean().item())
    # ---- all-pairs keyframe matching (BASELINE.json configs[4], scaled down): NCCL exchange of the
    # descriptor blocks, consumed block by block (distributed.allpairs_match_counts)
    allpairs = None
    ap_kf = args.allpairs_kf if args.allpairs_kf >= 0 else (4096 if world >= 8 else 512)   # config 5 in full on 8 GPUs
    if ap_kf > 0:
        from orb_slam2_detailed_comments_b200.distributed import NcclCommunicator, allpairs_match_counts_nccl, shard_range
        n_kf, n_desc = ap_kf, 1000
        rb, re = shard_range(n_kf, rank, world)
        g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
        local = torch.randint(0, 256, (re - rb, n_desc, 32), dtype=torch.uint8, device=dev, generator=g)
        all_desc = torch.empty((n_kf, n_desc, 32), dtype=torch.uint8, device=dev)
        ap_out = torch.empty((re - rb, n_kf), dtype=torch.int32, device=dev)
        comm = NcclCommunicator(local_rank) if world > 1 else None
        with torch.cuda.stream(tstream):
            allpairs_match_counts_nccl(matcher, comm, local, n_kf, all_desc, ap_out, stream=stream)   # warm-up
            barrier()
            a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
            a0.record(tstream)
            allpairs_match_counts_nccl(matcher, comm, local, n_kf, all_desc, ap_out, stream=stream)
            a1.record(tstream)
            torch.cuda.synchronize()
        ap_ms = max_over_ranks(a0.elapsed_time(a1))
        if comm is not None:
            comm.close()
        allpairs = {"keyframes": n_kf, "descriptors_per_keyframe": n_desc, "ms": ap_ms,
                    "value": float(n_kf) * n_kf * n_desc * n_desc / (ap_ms * 1e-3), "unit": "cmp/s", "scaling": "strong",
                    "api": "orb_match_allpairs_nccl (exchange inside the C ABI)",
                    "exchange": "ncclBroadcast per owner block on a communication stream, k_allpairs per landed block" if world > 1 else "none (1 GPU)",
                    "note": ("BASELINE config 5 at full size" if n_kf >= 4096 else
                             "BASELINE config 5 uses 4096 keyframes; scaled down to bound the run below 8 GPUs (--allpairs-kf 4096 runs it in full)")}
        del all_desc, ap_out, local
    # ---- BASELINE config 3: the EuRoC sequence, strong-sharded
    sequence = None
    seq_launches = 0
    if args.sequence_frames > 0:
        del d_imgs, d_kps, d_desc, d_dsc, d_m12
        torch.cuda.empty_cache()
        sequence, seq_launches = run_sequence_config3(args, rank, world, local_rank, dev, barrier, max_over_ranks)
    pipes = int_pipe_peak(local_rank)
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel + of the whole path ---------------------------------
    levels = []
    for l in range(8):
        levels.append((int(np.rint(np.float32(W) * ext.GetInverseScaleFactors()[l])), int(np.rint(np.float32(H) * ext.GetInverseScaleFactors()[l]))))
    ext(pool[0])
    mean_cand = float(sum(len(ext.stage_candidates(0, l)[0]) for l in range(8)))
    total_b, per_stage_b = algorithmic_bytes(levels, mean_kp, mean_cand)
    peak, peak_src = measured_hbm_peak()
    stage_ms = {k: v[0] for k, v in stages.items()}
    dom = max(stage_ms, key=stage_ms.get)
    n_launch = max(1, stages[dom][1])
    frames_total = frames_per_step * args.steps
    kernels_per_chunk = max(1, round(stages[dom][1] / max(1, stages["quadtree"][1])))   # pyramid: 1 + 7 + 1 border launches; FAST: one per level group
    frames_per_launch = frames_total / (n_launch / kernels_per_chunk)
    dom_ms = stage_ms[dom] / n_launch
    # bytes per launch / average launch duration == stage bytes over all frames / stage time
    achieved = per_stage_b[dom] * frames_total / (stage_ms[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": {"pyramid": "k_level0_border2+k_resize_strip+k_fill_borders", "fast": "k_fast_cells", "quadtree": "k_quadtree",
                                           "blur": "k_blur7", "describe": "k_describe_ring"}[dom],
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(dom),
                "peak_source": peak_src, "algorithmic_bytes_per_frame": per_stage_b[dom], "frames_per_launch": frames_per_launch,
                "avg_launch_ms": dom_ms, "share_of_step": stage_ms[dom] / max(1e-9, sum(stage_ms.values()))}
    path_gbs = (fps / world) * total_b / 1e9

    # ---- CPU baseline beside it (rank 0, N = 1 only) -----------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import orb_oracle as O
        cores = O.hardware_threads()
        sample = min(1024, max(64, 64 * cores))
        frames = np.concatenate([pool] * ((sample + len(pool) - 1) // len(pool)))[:sample]
        kind, what, cpu_extract = cpu_extractor_arm()
        t0 = time.perf_counter()
        cpu_extract(frames, cores)
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        O.extract_batch_mt(frames[: sample // 2], NFEAT, nthreads=cores)
        dt_port = (time.perf_counter() - t0) * 2
        pairs = max(8, 2 * cores)
        t1 = time.perf_counter()
        O.match_batch_mt(dsc[: 2 * min(pairs, uniq)], ang[: 2 * min(pairs, uniq)], 0.9, nthreads=cores)
        dtm = time.perf_counter() - t1
        t2 = time.perf_counter()
        for sc_ in tr_uniq[:4]:
            q_ = O.project_last_frame(sc_["Xw"], sc_["mp_flags"], sc_["last"], sc_["Tcw"], sc_["cam4"], sc_["bounds"], sc_["mbf"], 15.0, sfs, 0)
            O.search_by_projection(sc_["cur"], sc_["cur_desc"], sc_["uright"], sc_["bounds"], sc_["occupied0"], q_, sc_["mp_desc"], 0, 100, 0.9, True)
        tracking["cpu_oracle_ms_per_frame_1_thread"] = (time.perf_counter() - t2) / 4 * 1e3
        cv2_orb = opencv_orb_baseline(frames[: max(64, 16 * cores)], cores)
        cpu = {"value": sample / dt, "unit": "frames/s", "cores": cores, "kind": kind, "opencv_orb": cv2_orb,
               "sample": "%d %s-shape frames, one frame per thread; %s" % (sample, args.workload, what),
               "oracle_port_frames_per_s": sample / dt_port,
               "matching_cmp_per_s": min(pairs, uniq) * MATCH_N * MATCH_N / dtm}

    popc_peak_cmp = pipes["popc"] / 8.0
    line = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "%s-shape %dx%d stereo pairs, ORBextractor(%d,1.2,8,20,7) per eye, " % (args.workload, W, H, NFEAT) +
                               "%d pairs (%d eye-frames) per step per GPU" % (args.pairs, frames_per_step),
                   "frames_per_step_per_gpu": frames_per_step, "unique_frames": UNIQUE_FRAMES, "chunk_frames": args.chunk,
                   "l2": "inputs larger than L2 (%d MB per step; unique pool %d MB)" % (frames_per_step * W * H // 1000000, UNIQUE_FRAMES * W * H // 1000000), "parallelism": "frames sharded, no collective",
                   "numa_node_of_rank0": numa},
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "api": "orb_extract_batch_host_async per step + orb_synchronize at the end (pinned host buffers)",
                "value_blocking_calls": e2e_sync_fps, "api_blocking": "orb_extract_batch_host (returns when the step's results are on the host)",
                "h2d_GBps_in_run": e2e_fps / world * W * H / 1e9,
                "d2h_GBps_in_run": e2e_fps / world * (cap * 60 + 4) / 1e9,
                "h2d_GBps_copy_engine_alone": h2d_peak,
                "ceiling_frames_per_s": h2d_peak * 1e9 / (W * H) * world,
                "ceiling_frames_per_s_upload_and_download": bidir_fps,
                "frac_of_bidirectional_copy_ceiling": e2e_fps / bidir_fps,
                "limiter": "host <-> device copies: the copy engines alone (upload of the frames and download of the keypoints / descriptors "
                           "of a step, concurrently, all ranks at once) move a step's bytes at ceiling_frames_per_s_upload_and_download",
                "pinned_buffer_numa_node": pinned_node},
        "gpu_launches": launches_per_step * args.steps + e2e_launches + stereo_launches + frame_launches + track_launches + input_launches + m_steps + (2 * world if allpairs else 0) + seq_launches,
        "stereo": stereo,
        "frame_helpers": frame_helpers,
        "tracking": tracking,
        "input_stage": input_stage,
        "latency": lat,
        "clocks": clocks,
        "roofline": roofline,
        "path_roofline": {"bound": "hbm", "algorithmic_bytes_per_frame": total_b, "achieved": path_gbs, "peak": peak, "unit": "GB/s",
                          "frac": path_gbs / peak},
        "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
        "issue_roofline": issue_roofline(fps / world, clocks),
        "stage_roofline_frac": {k: per_stage_b[k] * frames_total / (max(1e-9, stage_ms[k]) * 1e-3) / 1e9 / peak for k in stage_ms},
        "stage_timing": {"how": "second pass of the same steps, chunks serialised on one stream, CUDA events between stages",
                         "ms_per_step_serialised": ms_serial / args.steps},
        "mean_keypoints": mean_kp, "mean_candidates": mean_cand,
        "matching": {"metric": "Hamming cmp/s (2000x2000 brute-force 2-NN, ratio 0.9, dedup, rotation histogram)",
                     "value": cmp_per_s, "unit": "cmp/s", "pairs_per_step": match_pairs, "steps": m_steps, "ms_per_step": m_ms / m_steps,
                     "mean_matches_per_pair": mean_matches,
                     "roofline": {"bound": "int-pipe", "achieved": cmp_per_s / world, "peak": popc_peak_cmp, "unit": "cmp/s",
                                  "frac": cmp_per_s / world / popc_peak_cmp,
                                  "peak_def": "measured POPC issue rate / 8 POPC per naive 256-bit comparison (SURVEY 8d)",
                                  "popc_ops_per_s": pipes["popc"], "lop3_ops_per_s": pipes["lop3"],
                                  "mix_units_per_s": pipes["mix_popc_4lop3"],
                                  "frac_of_mix_peak": cmp_per_s / world / (pipes["mix_popc_4lop3"] / 4.0),
                                  "mix_peak_def": "kernel issues 4 POPC + 16 LOP3 per comparison (carry-save adders); "
                                                  "peak = measured rate of that mix (POPC and LOP3 share the ALU pipe) / 4"}},
    }
    mix_peak_cmp = pipes["mix_popc_4lop3"] / 4.0
    line["roofline_matching"] = {"bound": "int-pipe", "kernel": "k_match_pairs_bf", "achieved": cmp_per_s / world, "peak": mix_peak_cmp,
                                 "unit": "cmp/s", "frac": cmp_per_s / world / mix_peak_cmp,
                                 "peak_def": "the kernel issues 4 POPC + 16 LOP3 per 256-bit comparison (carry-save adders); peak = the "
                                             "measured issue rate of that instruction mix on this GPU (orb_int_pipe_peak) / 4"}
    if allpairs:
        allpairs["roofline_frac"] = allpairs["value"] / world / mix_peak_cmp
        allpairs["roofline_frac_of_popc_over_8"] = allpairs["value"] / world / popc_peak_cmp
        line["allpairs"] = allpairs
    if sequence:
        sequence["path_roofline_frac"] = None
        line["sequence_config3"] = sequence
    if parity:
        line["parity"] = parity
    if cpu:
        line["cpu_baseline"] = cpu
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("ORB_CHUNK", "256")), help="frames per internal chunk")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--match-steps", type=int, default=3)
    ap.add_argument("--workload", default="kitti", choices=sorted(WORKLOADS), help="frame shape / feature count (default: the headline KITTI shape)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-call latency loops (profiling runs)")
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_STEP, help="stereo pairs per step per GPU (profiling runs shrink this)")
    ap.add_argument("--match-pairs", type=int, default=MATCH_PAIRS)
    ap.add_argument("--allpairs-kf", type=int, default=-1,
                    help="keyframes of the all-pairs workload (0 = skip; default: 4096 = BASELINE config 5 in full on 8 GPUs, 512 below)")
    ap.add_argument("--sequence-frames", type=int, default=8192, help="frames of the EuRoC sequence of BASELINE config 3, sharded across the GPUs (0 = skip)")
    ap.add_argument("--parity-frames", type=int, default=8, help="frames re-checked against the CPU oracle outside the timed regions (0 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    global W, H, NFEAT, METRIC
    W, H, NFEAT = WORKLOADS[args.workload][:3]
    if args.workload != "kitti":
        METRIC = "ORB-extract frames/s @%dx%d %d feats" % (W, H, NFEAT)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: re-launch under torchrun, one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
