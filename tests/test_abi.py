"""CPU-side checks of the C ABI library: it loads, exports every symbol include/orb_b200.h
declares, its host-side entry point works, and compute entry points FAIL LOUDLY without a GPU
(there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "orb_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(orb_[a-z0-9_]+)\s*\(", txt)))


def test_library_is_built_and_exports_every_declared_symbol():
    from orb_slam2_detailed_comments_b200 import _lib
    L = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "liborb_b200.so does not export %s" % n
    assert sorted(_lib.EXPORTS) == names


def test_keypoint_record_is_cv_keypoint_sized():
    from orb_slam2_detailed_comments_b200 import KP_DTYPE
    assert KP_DTYPE.itemsize == 28
    assert [KP_DTYPE.fields[f][1] for f in ("x", "y", "size", "angle", "response", "octave", "class_id")] == [0, 4, 8, 12, 16, 20, 24]


def test_descriptor_distance_host(oracle):
    from orb_slam2_detailed_comments_b200 import ORBmatcher
    rng = np.random.RandomState(0)
    for _ in range(200):
        a = rng.randint(0, 256, 32).astype(np.uint8); b = rng.randint(0, 256, 32).astype(np.uint8)
        assert ORBmatcher.DescriptorDistance(a, b) == oracle.hamming(a, b) == int(np.unpackbits(a ^ b).sum())


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from orb_slam2_detailed_comments_b200 import ORBextractor, ORBmatcher, OrbError
    with pytest.raises(OrbError) as e:
        ORBextractor(1000, 1.2, 8, 20, 7)
    assert e.value.status == 2  # ORB_ERR_CUDA
    with pytest.raises(OrbError):
        ORBmatcher(0.9, True)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "orb_slam2_detailed_comments_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower() or f == "synth.py" and False, "%s mentions the oracle" % f


def test_dropin_reference_build_loads_and_fails_loudly_without_gpu():
    """oracle/_ref/liborbref_gpu.so = the reference's unmodified Frame.cc / ORBmatcher.cc + the drop-in translation units
    (compat/orb_b200_*.cpp, compiled against the reference's unmodified headers) + liborb_b200.so."""
    import torch
    from oracle import orb_refgpu
    if not orb_refgpu.available():
        orb_refgpu.build()
    if not orb_refgpu.available():
        pytest.skip("needs /root/reference to build")
    L = orb_refgpu.lib()
    for n in orb_refgpu.EXPORTS:
        assert hasattr(L, n), n
    a = np.arange(32, dtype=np.uint8); b = a ^ 0x11
    assert orb_refgpu.descriptor_distance(a, b) == 64     # ORBmatcher::DescriptorDistance through the reference's class
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError) as e:
            orb_refgpu.DropInExtractor(1000, 1.2, 8, 20, 7)
        assert "liborb_b200" in str(e.value)
