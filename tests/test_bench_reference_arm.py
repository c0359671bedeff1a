"""The reference arm of bench.py (CPU oracle port, no GPU needed) prints exactly one JSON line with
the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--workload", "tum1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["higher_is_better"] is True


def test_rank_cpu_block_fallback_when_numa_is_hidden():
    """bench.bind_to_gpu_numa_node: without sysfs NUMA information every rank pins itself to its own contiguous block of the
    visible CPUs (checked in a subprocess so that this process keeps its affinity)."""
    code = r'''
import os, sys, types
sys.path.insert(0, %r)
import bench
avail = sorted(os.sched_getaffinity(0))
props = types.SimpleNamespace(pci_domain_id=0xffff, pci_bus_id=0xff, pci_device_id=0x1f)   # no such device in sysfs
got = bench.bind_to_gpu_numa_node(props)
now = sorted(os.sched_getaffinity(0))
world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
if len(avail) >= 2 * world:
    per = len(avail) // world
    assert now == avail[rank * per:(rank + 1) * per], (now, avail)
    assert isinstance(got, str) and got.startswith("cpu-block")
else:
    assert now == avail and got is None
print("ok")
''' % ROOT
    env = dict(os.environ, WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert out.returncode == 0 and "ok" in out.stdout + out.stderr, out.stdout + out.stderr   # bench keeps stdout for its JSON line
