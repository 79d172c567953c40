"""The decision pass of whole pictures on the device (xb200_analyze_picture) -- see tests/picture_on_device.py, which runs in a
subprocess so that a hang or a sticky CUDA error cannot touch the other GPU tests.  Hard assertions: coder states before / after
every CTU, frame maps, leaf CUs, pictures before / after deblocking equal the reference's, and the records injected into the
unmodified reference give the byte-identical bitstream."""
import os
import subprocess
import sys

import pytest

SCRIPT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "picture_on_device.py")


def _run(*args, timeout=900):
    r = subprocess.run([sys.executable, SCRIPT, *args], capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0 and "PICTURE_ON_DEVICE_OK" in r.stdout, r.stdout[-2500:] + r.stderr[-2500:]
    print(r.stdout)
    return r.stdout


@pytest.mark.gpu
def test_fixture_and_live_default_gop_threads_1_and_2():
    """committed fixture (3 pictures) + a live QCIF encode of the default hierarchical-B GOP with one and two coder-state chains"""
    out = _run()
    assert out.count("byte-identical bitstream") == 2


@pytest.mark.gpu
def test_other_configurations():
    """10-bit input preset medium, plain quantiser at lower QP, low QP with three row chains, the whole default GOP + 3 pictures"""
    out = _run("--more")
    assert out.count("byte-identical bitstream") == 4


@pytest.mark.gpu
def test_full_size_1080p_default_gop_eight_chains():
    """BASELINE.json configs[1] size: 1920x1080 8-bit preset fast, the first pictures of the default GOP (I + two B pictures with two-sided
    references), decided as eight coder-state chains per picture like the reference run with -m 8"""
    out = _run("--config", "1080p:fast:3:8")
    assert "byte-identical bitstream" in out


@pytest.mark.gpu
def test_2160p_ten_bit_medium_eight_chains():
    """BASELINE.json configs[2] size: 3840x2160 10-bit preset medium, I + one B picture, eight chains"""
    out = _run("--config", "2160p10:medium:2:8", timeout=1500)
    assert "byte-identical bitstream" in out
