#!/usr/bin/env python
"""Mint tests/golden/*.npz from the cv2-driven restatement of the reference (tests/cv2_reference.py).

Runs in the dev container (needs python cv2 4.13.0; does NOT read /root/reference). The
fixtures pin the oracle and, through it, the CUDA path: input image, keypoints, descriptors,
per-level CRC32 of the bordered pyramid and the blurred levels, candidate and kept counts.
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cv2_reference import Cv2Reference, load_pattern  # noqa: E402
from orb_slam2_detailed_comments_b200.synth import adversarial_frames, synth_frame  # noqa: E402


def golden_for(img, nfeat):
    ref = Cv2Reference(nfeat, 1.2, 8, 20, 7, load_pattern())
    kps, desc, levels, dbg = ref(img)
    kps = np.asarray(kps, np.float64).reshape(-1, 6)
    out = dict(
        image=img, nfeatures=np.int32(nfeat),
        kp_xy=kps[:, 0:2].astype(np.float32), kp_size=kps[:, 2].astype(np.float32),
        kp_angle=kps[:, 3].astype(np.float32), kp_response=kps[:, 4].astype(np.float32),
        kp_octave=kps[:, 5].astype(np.int32), descriptors=desc.astype(np.uint8),
        level_crc=np.array([zlib.crc32(np.ascontiguousarray(l).tobytes()) for l in levels], np.uint32),
        blur_crc=np.array([zlib.crc32(np.ascontiguousarray(d["blur"]).tobytes()) if "blur" in d else 0 for d in dbg], np.uint32),
        n_candidates=np.array([len(d["cand"]) for d in dbg], np.int32),
        n_kept=np.array([len(d["kept"]) for d in dbg], np.int32),
        fallback_cells=np.array([d["fallback"] for d in dbg], np.int32),
    )
    return out


def main():
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    np.savez_compressed(os.path.join(gold, "synth_320x240_f400.npz"), **golden_for(synth_frame(320, 240, 1), 400))
    np.savez_compressed(os.path.join(gold, "synth_401x257_f500.npz"), **golden_for(synth_frame(401, 257, 2), 500))
    adv = adversarial_frames(320, 240)
    for k in ("low_contrast", "checkerboard", "uniform_noise", "constant"):
        np.savez_compressed(os.path.join(gold, "adv_%s_320x240_f300.npz" % k), **golden_for(adv[k], 300))


if __name__ == "__main__":
    main()
