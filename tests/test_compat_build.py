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
def test_adapter_runs_on_gpu(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe, "run"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "keypoints" in out.stdout and "pyramid0 640x480" in out.stdout
    print(out.stdout.strip().splitlines()[-1])   # adapter latency line
