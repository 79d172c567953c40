"""bench.py, host side: the reference arm (`--impl reference`) runs entirely on the host cores -- the same application-level program
as the device arm (integration/xb200_streams.c) linked against the unmodified reference library -- so its JSON line can be checked
without a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


@pytest.mark.skipif(not os.path.exists(os.path.join(REFDIR, "xeveb_streams_ref")), reason="oracle/_ref not built here")
def test_reference_arm_emits_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cif", "--streams", "2", "--frames", "9",
                        "--steps", "2", "--warmup", "1", "--threads", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "encoded pictures/s" and d["unit"] == "pictures/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 2 and cb["value"] == d["value"] and cb["md5_unique"] is True and len(cb["md5"]) == 32
    assert d["e2e"] == {"value": d["value"], "unit": "pictures/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "352x288" in d["config"]["workload"] and "threads=2" in d["config"]["workload"]


def test_usable_cores_respects_affinity_and_quota():
    sys.path.insert(0, ROOT)
    import bench
    n = bench.usable_cores()
    assert 1 <= n <= (os.cpu_count() or 1)
