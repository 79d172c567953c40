"""Main profile, first device operator: xb200_transform_main -- see tests/transform_main_on_device.py."""
import os
import subprocess
import sys

import pytest

SCRIPT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "transform_main_on_device.py")


@pytest.mark.gpu
def test_main_profile_transforms_match_oracle():
    """IQT DCT-II (2..64 points) and ATS (DST-VII / DCT-VIII, 4..32 points), forward and inverse, square and non-square blocks, 8 and
    10 bit, bit-exact against the oracle that is pinned to the Main-profile reference; invalid items are rejected.
    Runs in a subprocess (nothing it does to the CUDA context can touch the other tests).  First hardware run: round 1's driver box
    (GPUTEST_r01: passed); a hard assertion since."""
    r = subprocess.run([sys.executable, SCRIPT], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "TRANSFORM_MAIN_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
    print(r.stdout)
