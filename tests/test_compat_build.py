"""The C++ adapter (compat/orb_b200_compat.hpp: the reference's ORBextractor / ORBmatcher
signatures over the C ABI) compiles with plain g++, links against liborb_b200.so and, on a GPU,
extracts and matches through the reference-shaped calls."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "orb_slam2_detailed_comments_b200", "lib")


def _build(tmp_path):
    exe = str(tmp_path / "compat_probe")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "compat_probe.cpp"),
                           "-L" + LIBDIR, "-lorb_b200", "-Wl,-rpath," + LIBDIR])
    return exe


def test_adapter_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "distance 64" in out.stdout


@pytest.mark.gpu
def test_adapter_runs_on_gpu(tmp_path, oracle):
    """Through the header-only C++ adapter: the keypoints, descriptors and SearchForInitialization matches it returns equal the
    oracle's on the same image (the probe's LCG checkerboard, regenerated here)."""
    import numpy as np
    from orb_slam2_detailed_comments_b200._lib import KP_DTYPE
    exe = _build(tmp_path)
    dump = str(tmp_path / "adapter.bin")
    out = subprocess.run([exe, "run", dump], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "keypoints" in out.stdout and "pyramid0 640x480" in out.stdout
    print(out.stdout.strip().splitlines()[-1])   # adapter latency line
    W, H = 640, 480
    lcg = np.zeros(W * H, np.uint32)
    v = 12345
    for i in range(W * H):   # s = s * 1664525 + 1013904223 (mod 2^32), row-major
        v = (v * 1664525 + 1013904223) & 0xffffffff
        lcg[i] = v
    yy, xx = np.mgrid[0:H, 0:W]
    img = ((((xx // 24 + yy // 24) & 1) * 120 + 60) + ((lcg.reshape(H, W) >> 24) & 15)).astype(np.uint8)
    raw = open(dump, "rb").read()
    nk = int(np.frombuffer(raw, np.int32, 1, 0)[0])
    kps = np.frombuffer(raw, KP_DTYPE, nk, 4)
    desc = np.frombuffer(raw, np.uint8, nk * 32, 4 + 28 * nk).reshape(nk, 32)
    off = 4 + 60 * nk
    n, nm = [int(x) for x in np.frombuffer(raw, np.int32, 2, off)]
    m12 = np.frombuffer(raw, np.int32, nm, off + 8)
    okps, odesc = oracle.OracleExtractor(1000, 1.2, 8, 20, 7)(img)
    assert nk == len(okps) > 500
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(kps[f], okps[f]), f
    d = np.abs(kps["angle"] - okps["angle"])
    assert np.minimum(d, 360.0 - d).max() <= 1e-3
    assert (desc == odesc).all(1).mean() >= 0.999
    xy = np.stack([okps["x"], okps["y"]], 1).astype(np.float32)
    n_ref, m_ref, _, _, _ = oracle.search_for_initialization(xy, okps["octave"], okps["angle"], odesc, xy, okps["octave"], okps["angle"], odesc,
                                                             (0, W, 0, H), xy.copy(), window=100, nnratio=0.9, check_ori=True, mode=0)
    assert n == n_ref and np.array_equal(m12, m_ref)
