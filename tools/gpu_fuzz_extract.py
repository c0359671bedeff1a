#!/usr/bin/env python
"""One-off stress of the batch path (chunks of >= 8 frames: k_describe_ring) on random image sizes, feature budgets, scale factors
and level counts against the CPU oracle. Usage (GPU box): python tools/gpu_fuzz_extract.py [n_configs] [seed]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import orb_oracle as O  # noqa: E402
from orb_slam2_detailed_comments_b200 import KP_DTYPE, ORBextractor  # noqa: E402
from orb_slam2_detailed_comments_b200.synth import synth_batch  # noqa: E402

n_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
rows = flips = 0
for it in range(n_cfg):
    w = int(rng.randint(160, 1400)); h = int(rng.randint(120, 900))
    nf = int(rng.choice([50, 300, 1000, 2000, 3500])); sf = float(rng.choice([1.1, 1.2, 1.25, 1.33])); nl = int(rng.randint(2, 9))
    ini = int(rng.choice([20, 30, 12])); mn = int(rng.choice([7, 5, 12]))
    B = int(rng.randint(8, 14))
    try:
        gpu = ORBextractor(nf, sf, nl, ini, mn, max_batch=16)
    except Exception as e:
        print("config", it, (w, h, nf, sf, nl), "create:", str(e)[:80]); continue
    imgs = synth_batch(w, h, B, seed0=int(rng.randint(1 << 20)))
    try:
        cap = gpu.max_keypoints_for(w, h)
        if cap <= 0:
            print("config", it, (w, h, nf, sf, nl), "unsupported geometry"); continue
        d_k = torch.zeros((B, cap, 28), dtype=torch.uint8, device="cuda"); d_d = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
        d_c = torch.zeros(B, dtype=torch.int32, device="cuda")
        gpu.extract_batch_device(torch.from_numpy(imgs).cuda(), d_k, d_d, d_c)
        gpu.synchronize()
    except Exception as e:
        print("config", it, (w, h, nf, sf, nl), "extract:", str(e)[:100]); continue
    c = d_c.cpu().numpy(); k = d_k.cpu().numpy().view(KP_DTYPE).reshape(B, cap); d = d_d.cpu().numpy()
    orc = O.OracleExtractor(nf, sf, nl, ini, mn)
    ok = True
    for b in (0, B - 1):
        okps, odesc = orc(imgs[b])
        n = c[b]
        if n != len(okps):
            ok = False; print("  count", n, len(okps)); continue
        for f in ("x", "y", "octave", "response", "size"):
            if not np.array_equal(k[b, :n][f], okps[f]):
                ok = False; print("  field", f)
        if n:
            da = np.abs(k[b, :n]["angle"] - okps["angle"]); da = np.minimum(da, 360 - da)
            if da.max() > 1e-3:
                ok = False; print("  angle", da.max())
            diff = int((~(d[b, :n] == odesc).all(1)).sum())
            rows += n; flips += diff
            if diff > 0.001 * n:
                ok = False; print("  desc rows differ", diff, "of", n)
    print("config", it, (w, h, nf, sf, nl, ini, mn, B), "keypoints", int(c[0]), "OK" if ok else "MISMATCH", flush=True)
    bad += not ok
    gpu.close()
print("configs with mismatches:", bad, "descriptor rows compared", rows, "differing", flips)
sys.exit(1 if bad else 0)
