#!/usr/bin/env python
"""Runs every kernel of the SURVEY 8(f) "next" rows twice on bench-sized inputs, for `ncu --set full` captures
(tools/gpu_evidence.sh): input stage, frame helpers, tracking matchers, distinctive descriptors."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_detailed_comments_b200 import KP_DTYPE, frame as F, input as IN, search as SR  # noqa: E402
from orb_slam2_detailed_comments_b200.synth import tracking_scene  # noqa: E402

dev = torch.device("cuda", 0)
W, H, B, n = 1241, 376, 256, 2000
rng = np.random.RandomState(0)
d_rgb = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev)
d_g = torch.zeros((B, H, W), dtype=torch.uint8, device=dev)
yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=dev), torch.arange(W, dtype=torch.float32, device=dev), indexing="ij")
rr = ((xx - W / 2) ** 2 + (yy - H / 2) ** 2) / float(W * W)
d_mx = (xx + (xx - W / 2) * 0.08 * rr + 1.3).contiguous(); d_my = (yy + (yy - H / 2) * 0.08 * rr - 0.7).contiguous()
d_rect = torch.zeros_like(d_g)
counts = rng.randint(2, 25, 100000).astype(np.int32)
d_off = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)).to(dev)
d_desc = torch.randint(0, 256, (int(counts.sum()), 32), dtype=torch.uint8, device=dev)
d_best = torch.zeros(len(counts), dtype=torch.int32, device=dev)

sc = [tracking_scene(n, n, 4242 + i, w=W, h=H, distinct=0.97) for i in range(8)]
SF = np.cumprod(np.concatenate([[1.0], np.full(7, np.float32(1.2), np.float32)]).astype(np.float32)).astype(np.float32)


def tile(key, kp=False):
    a = np.stack([s[key].view(np.uint8).reshape(n, 28) if kp else np.ascontiguousarray(s[key]) for s in sc])
    return torch.from_numpy(a).to(dev).repeat((B // len(sc),) + (1,) * (a.ndim - 1)).contiguous()


t_kps, t_desc, t_ur, t_occ = tile("cur", True), tile("cur_desc"), tile("uright"), tile("occupied0")
t_last, t_Xw, t_fl, t_mpd = tile("last", True), tile("Xw"), tile("mp_flags"), tile("mp_desc")
t_T = torch.from_numpy(np.stack([s["Tcw"] for s in sc])).to(dev).repeat(B // len(sc), 1, 1).contiguous()
t_cnt = torch.full((B,), n, dtype=torch.int32, device=dev); t_dir = torch.zeros(B, dtype=torch.int32, device=dev)
t_cs = torch.zeros((B, 3073), dtype=torch.int32, device=dev); t_ci = torch.zeros((B, n), dtype=torch.int32, device=dev)
t_q = torch.zeros((B, n, 32), dtype=torch.uint8, device=dev); t_un = torch.zeros_like(t_kps)
t_mk = torch.zeros((B, n), dtype=torch.int32, device=dev); t_mq = torch.zeros((B, n), dtype=torch.int32, device=dev)
t_nm = torch.zeros(B, dtype=torch.int32, device=dev)
t_scr = torch.zeros(SR.scratch_bytes(B, n, n), dtype=torch.uint8, device=dev)
tb = sc[0]["bounds"]
cam = F.camera(718.856, 718.856, 607.19, 185.2, -0.2834, 0.0739, 0.00019, 1.76e-05, 0.0)
node2 = torch.randint(0, 400, (B, n), dtype=torch.int32, device=dev)
node1 = node2.clone()
us = (t_fl & 1).contiguous()

for _ in range(2):
    IN.cvtColorGray(d_rgb, IN.RGB2GRAY, d_g)
    IN.remap(d_g, d_mx, d_my, d_rect)
    IN.ComputeDistinctiveDescriptors(d_desc, d_off, 32, d_best)
    F.UndistortKeyPoints(t_kps, t_cnt, cam, t_un)
    F.AssignFeaturesToGrid(t_kps, t_cnt, tb, t_cs, t_ci)
    SR.ProjectLastFrame(t_Xw, t_fl, t_last, t_cnt, t_T, t_dir, sc[0]["cam4"], tb, sc[0]["mbf"], 15.0, SF, t_q)
    fr = SR.device_frames(t_kps, t_desc, t_cnt, tb, t_cs, t_ci, t_ur, t_occ)
    SR.SearchByProjection(fr, t_q, t_mpd, t_cnt, SR.ORB_SEARCH_BEST, SR.TH_HIGH, 0.9, True, t_scr, t_mk, t_mq, t_nm)
    SR.SearchByBoW(t_last, t_mpd, node1, us, t_cnt, fr, node2, 0.7, True, t_scr, t_mk, t_mq, t_nm)
torch.cuda.synchronize()
print("matches per frame", float(t_nm.float().mean()))
